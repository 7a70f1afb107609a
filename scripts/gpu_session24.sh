#!/bin/bash
# A/B on one box: previous library (walk-only entries) vs current (variant entries); block sizes with variants
mkdir -p gpurun_out
nvidia-smi -L
echo "== previous library"; GMB_LIB_PATH=$PWD/genmap_b200/lib/variants/libgenmap_b200_prev.so timeout 600 python tools/sweep.py --reps 3 --configs 0:-1:256,1:-1:64,2:-1:8 > gpurun_out/s24_sweep_prev.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/s24_sweep_prev.log
echo "== current library"; timeout 900 python tools/sweep.py --reps 3 --configs 0:-1:256,1:-1:64,2:-1:8,1:-1:64:2,1:-1:64:3,1:-1:64:4,1:-1:64:5,1:-1:64:6,1:-1:64:8,2:-1:8:2,2:-1:8:3,2:-1:8:4,2:-1:8:5,2:-1:8:6 > gpurun_out/s24_sweep_cur.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/s24_sweep_cur.log
echo "== parity subset"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "oracle or golden or fixtures" 2>&1 | tail -3
