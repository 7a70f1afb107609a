#!/bin/bash
# resident CTAs per SM for the blocked Dna4 kernels now that they are issue-bound: 3 / 4 (default) / 5
mkdir -p gpurun_out
for v in mb3 "" mb5; do
  echo "== variant '$v'"
  if [ -n "$v" ]; then export GMB_LIB_PATH=$PWD/genmap_b200/lib/variants/libgenmap_b200_$v.so; else unset GMB_LIB_PATH; fi
  timeout 600 python tools/sweep.py --reps 3 --configs 1:-1:64,2:-1:8,3:-1:0.5 > gpurun_out/s28_sweep_$v.log 2>&1; grep -v "fetches by\|genome" gpurun_out/s28_sweep_$v.log
done
