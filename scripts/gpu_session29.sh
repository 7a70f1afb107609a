#!/bin/bash
# 3 CTAs per SM for the blocked Dna4 kernels up to E = 2: parity subset + CLI golden replay + sweep
mkdir -p gpurun_out
echo "== parity subset"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "oracle or golden or fixtures or block_size or jump or pangenome or 64 or edge" 2>&1 | tail -3
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== sweep"; timeout 600 python tools/sweep.py --reps 3 --configs 0:-1:256,1:-1:64,2:-1:8,3:-1:0.5 > gpurun_out/s29_sweep.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/s29_sweep.log
