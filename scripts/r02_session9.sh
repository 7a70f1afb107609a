#!/bin/bash
# round 2, session 9: after reverting the shared walk phase — kernel parity, default sweep, block kernel at E = 3, 4
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest kernels"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_s9_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r02_s9_pytest.log
echo "== default"; timeout 600 python tools/sweep.py --reps 3 --configs 0:-1:256,1:-1:64,2:-1:8,3:-1:0.5,4:-1:0.03125 2>&1 | tee gpurun_out/r02_s9_sweep_default.log | grep -v "fetches by"
echo "== block kernel forced for E = 3, 4"; GMB_BLOCK_KERNEL=2 timeout 600 python tools/sweep.py --reps 3 --configs 3:-1:0.5,4:-1:0.03125,4:-1:0.125 2>&1 | tee gpurun_out/r02_s9_sweep_block_e34.log | grep -v "fetches by"
