#!/bin/bash
# round 2, session 2: located table entries (keys that occur once end their search at the table read) —
# parity suite of the kernels, A/B against tables without them (GMB_LOCATE=0) at 3 Gbp, block sizes with them
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest kernels"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_s2_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02_s2_pytest.log
echo "== sweep without located entries"; GMB_LOCATE=0 timeout 600 python tools/sweep.py --reps 3 --configs 0:-1:256,1:-1:64,2:-1:8,3:-1:0.5 > gpurun_out/r02_s2_sweep_locate0.log 2>&1; echo "rc=$?"; cat gpurun_out/r02_s2_sweep_locate0.log
echo "== sweep with located entries"; timeout 600 python tools/sweep.py --reps 3 --configs 0:-1:256,1:-1:64,2:-1:8,3:-1:0.5 > gpurun_out/r02_s2_sweep_locate1.log 2>&1; echo "rc=$?"; cat gpurun_out/r02_s2_sweep_locate1.log
echo "== block sizes with located entries"; timeout 900 python tools/sweep.py --reps 2 --configs 1:-1:64:3,1:-1:64:4,1:-1:64:6,1:-1:64:8,1:-1:64:10,2:-1:8:3,2:-1:8:4,2:-1:8:5,2:-1:8:6,2:-1:8:8 > gpurun_out/r02_s2_sweep_blocks.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/r02_s2_sweep_blocks.log
