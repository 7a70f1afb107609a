#!/bin/bash
# round 2, session 12: table entries of keys that occur twice carry both positions ("located pairs") — kernel parity, sweeps
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest kernels"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_s12_pytest_kernels.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r02_s12_pytest_kernels.log
echo "== default"; timeout 600 python tools/sweep.py --reps 3 --configs 0:-1:256,1:-1:64,2:-1:8,3:-1:0.5,4:-1:0.125 2>&1 | tee gpurun_out/r02_s12_sweep_pairs.log | grep -v "fetches by"
echo "== block sizes"; timeout 600 python tools/sweep.py --reps 2 --configs 1:-1:64:6,1:-1:64:8,1:-1:64:10,2:-1:8:3,2:-1:8:4,2:-1:8:5 2>&1 | tee gpurun_out/r02_s12_sweep_pairs_blocks.log | grep -v "fetches by"
echo "== K=50"; timeout 600 python tools/sweep.py --kmer 50 --reps 2 --configs 0:-1:256,2:-1:8 2>&1 | tee gpurun_out/r02_s12_sweep_pairs_k50.log | grep -v "fetches by"
