#!/bin/bash
# profiles/r02_v5: GPU tests, bench lines (c2 + reference arm, c3, real, c4), launch list of c2, one full ncu capture of the
# kernels that changed (sa_place, the digit pass from the text, the text histogram), sanitizer passes, fuzz.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
scripts/gpu_check.sh tests smoke bench
timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 exit $?"
timeout 600 python bench.py --workload real --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_real.json 2> gpurun_out/bench_real.err; echo "bench real exit $?"
timeout 900 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 exit $?"
for w in c2 real; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$w.csv \
    python bench.py --workload $w --steps 1 --warmup 1 --no-cpu > gpurun_out/launches_${w}_run.log 2>&1; echo "ncu launches $w exit $?"
done
for k in sa_place_kernel rs_hist_text_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_$k \
      python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
# the first digit pass (keys from the text) is launch 0, 3, 6 ... of rs_pass_kernel; a plain pass follows it
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_pass_kernel -s 12 -c 2 -f -o gpurun_out/prof_rs_pass_kernel \
    python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_rs_pass_kernel.log 2>&1; echo "ncu rs_pass exit $?"
scripts/gpu_check.sh sanitize
timeout 500 python scripts/gpu_fuzz.py ${FUZZ_SECONDS:-60} ${FUZZ_SEED:-3} > gpurun_out/fuzz_${FUZZ_SEED:-3}.json 2> gpurun_out/fuzz_${FUZZ_SEED:-3}.err; echo "fuzz exit $?"; cat gpurun_out/fuzz_${FUZZ_SEED:-3}.json
for w in c2 c3 real c4; do python scripts/bench_brief.py gpurun_out/bench_$w.json; done
