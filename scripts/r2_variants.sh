#!/bin/bash
# tuning run: every build under build/variants on the given workloads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in ${WORKLOADS:-c2 real}; do
  for so in build/variants/*.so; do
    timeout 300 python scripts/variant_bench.py $so $w 5 2>/dev/null | tail -1
  done
done | tee gpurun_out/variants.jsonl
