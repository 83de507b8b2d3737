#!/bin/bash
# ncu --set full of the recursion kernels (one capture each) during a 2 x 1 Mbp rem run, and of the sweep kernel during a bench step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in small_step_kernel split_apply_kernel bubble_apply_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s ${NCU_SKIP:-5} -c 1 -f -o gpurun_out/prof_$k \
      python scripts/rem_bench.py 2 1000000 ours > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
for k in pair_count_kernel sa_lead_kernel rs_pass_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_$k \
      python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
