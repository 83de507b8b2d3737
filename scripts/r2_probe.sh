#!/bin/bash
# Round-2 probe: full GPU parity suite, then the repeat-bearing inputs (real fixture, synthetic repeats, graph-like text):
# a bench line each, and the ncu launch lists of the real-data and repeats steps.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^real|passed|failed|Error|error|exit" gpurun_out/pytest_gpu.log | tail -25
for w in ${WORKLOADS:-real repeats graph c2 c3}; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w exit $?"
  tail -2 gpurun_out/bench_$w.err; python scripts/bench_brief.py gpurun_out/bench_$w.json
done
for w in ${NCU_WORKLOADS:-real repeats}; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$w.csv \
    python bench.py --workload $w --steps 1 --warmup 1 --no-cpu > gpurun_out/launches_${w}_run.log 2>&1; echo "ncu launches $w exit $?"
done
if [ -n "$NCU_KERNEL" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$NCU_KERNEL -s ${NCU_SKIP:-2} -c ${NCU_COUNT:-1} -f -o gpurun_out/prof_$NCU_KERNEL \
      python bench.py --workload ${NCU_WORKLOAD:-c2} --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_run.log 2>&1; echo "ncu full exit $?"
fi
