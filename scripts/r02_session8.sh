#!/bin/bash
# round 2, session 8: warp-shared walk phase of block_kernel.cu — kernel parity suite (incl. -ep), sweeps for E = 1..4
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest kernels"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_s8_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02_s8_pytest.log
echo "== block kernel with the shared walk phase"; timeout 600 python tools/sweep.py --reps 3 --configs 1:-1:64,2:-1:8,2:-1:8:3 2>&1 | tee gpurun_out/r02_s8_sweep_shared.log | grep -v "fetches by"
echo "== the same for E = 3, 4 (GMB_BLOCK_KERNEL=2) against the general kernel"; GMB_BLOCK_KERNEL=2 timeout 600 python tools/sweep.py --reps 3 --configs 3:-1:0.5,4:-1:0.03125 2>&1 | tee gpurun_out/r02_s8_sweep_shared_e34.log | grep -v "fetches by"
timeout 600 python tools/sweep.py --reps 3 --configs 3:-1:0.5,4:-1:0.03125 2>&1 | tee gpurun_out/r02_s8_sweep_general_e34.log | grep -v "fetches by"
