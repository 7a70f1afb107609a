#!/bin/bash
# round 2, session 11: round size / put-aside slots of block_kernel.cu; the three kernels and both table flavours agree
mkdir -p gpurun_out
nvidia-smi -L
for v in round64 slots12; do
  echo "-- $v"; GMB_LIB_PATH=$PWD/genmap_b200/lib/variants/libgenmap_b200_$v.so timeout 600 python tools/sweep.py --reps 3 --configs 1:-1:64,2:-1:8 2>&1 | grep -v "fetches by" | tee gpurun_out/r02_s11_sweep_$v.log
done
echo "-- default"; timeout 600 python tools/sweep.py --reps 3 --configs 1:-1:64,2:-1:8 2>&1 | grep -v "fetches by" | tee gpurun_out/r02_s11_sweep_default.log
echo "== kernels agree"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "three_kernels" 2>&1 | tail -5
