#!/bin/bash
# One gpurun call: GPU parity tests, a bench line, the ncu launch list and one
# full capture of the dominant kernel.  Outputs land in gpurun_out/.
# usage: scripts/gpu_check.sh [tests] [bench] [launches] [ncu]   (default: all)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what="${*:-tests bench launches ncu}"
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for w in $what; do
case $w in
tests)
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -15 gpurun_out/pytest_gpu.log ;;
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
  tail -5 gpurun_out/smoke.log ;;
bench)
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench exit $?"
  tail -3 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c2_ref.json 2> gpurun_out/bench_c2_ref.err
  cat gpurun_out/bench_c2_ref.json ;;
bench3)
  timeout 900 python bench.py --workload c3 --steps 3 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 exit $?"
  tail -3 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json ;;
bench4)
  timeout 1200 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 exit $?"
  tail -3 gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.json ;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/launches_run.log 2>&1; echo "ncu launches exit $?" ;;
ncu)
  # NCU_KERNEL (regex, default rs_pass_kernel), NCU_SKIP (launches of that kernel to skip), NCU_WORKLOAD
  k="${NCU_KERNEL:-rs_pass_kernel}"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s ${NCU_SKIP:-3} -c ${NCU_COUNT:-2} -f -o gpurun_out/prof_$k \
      python bench.py --workload ${NCU_WORKLOAD:-c2} --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_run.log 2>&1; echo "ncu full exit $?" ;;
sanitize)
  timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or tiny or ragged" > gpurun_out/memcheck.log 2>&1; echo "memcheck exit $?"
  tail -5 gpurun_out/memcheck.log
  timeout 1200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden" > gpurun_out/racecheck.log 2>&1; echo "racecheck exit $?"
  tail -5 gpurun_out/racecheck.log ;;
esac
done
