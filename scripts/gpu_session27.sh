#!/bin/bash
# early exit for empty table entries + table-read cost in the model: parity subset, sweep
mkdir -p gpurun_out
echo "== parity subset"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "oracle or golden or fixtures or block_size or jump or pangenome or 64" 2>&1 | tail -3
echo "== sweep"; timeout 900 python tools/sweep.py --reps 3 --configs 0:-1:256,1:-1:64,2:-1:8,3:-1:0.5,2:-1:8:3,2:-1:8:4,1:-1:64:6,1:-1:64:8 > gpurun_out/s27_sweep.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/s27_sweep.log
