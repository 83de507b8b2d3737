#!/bin/bash
# Round-2 profile set: launch lists (c2, real) and one full ncu capture per kernel that matters.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in c2 real c3; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$w.csv \
    python bench.py --workload $w --steps 1 --warmup 1 --no-cpu > gpurun_out/launches_${w}_run.log 2>&1; echo "ncu launches $w exit $?"
done
for k in sa_place_kernel sa_lead_kernel rs_pass_kernel pair_count_kernel sa_textprep_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_$k \
      python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:multi_count_kernel -s 3 -c 1 -f -o gpurun_out/prof_multi_count_kernel \
    python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_multi_count_kernel.log 2>&1; echo "ncu multi_count exit $?"
for k in lcp_sparse_kernel sa_apply_kernel sa_exact_gather_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_$k \
      python bench.py --workload real --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_dp_kernel -s 3 -c 1 -f -o gpurun_out/prof_chain_dp_kernel \
    python scripts/rem_bench.py 2 1000000 ours > gpurun_out/ncu_chain.log 2>&1; echo "ncu chain exit $?"
