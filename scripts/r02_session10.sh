#!/bin/bash
# round 2, session 10: clustered-read microbenchmark (would low-order substitutions make table reads cheaper?), E = 4 routing
mkdir -p gpurun_out
nvidia-smi -L
echo "== clusterread"; timeout 300 tools/clusterread 2>&1 | tee gpurun_out/r02_s10_clusterread.log
echo "== E=4 general vs block at the same batch"; GMB_BLOCK_KERNEL=0 timeout 600 python tools/sweep.py --reps 2 --configs 4:-1:0.125 2>&1 | grep -v "fetches by" | tee gpurun_out/r02_s10_e4_general.log
timeout 600 python tools/sweep.py --reps 2 --configs 4:-1:0.125,3:-1:0.5 2>&1 | grep -v "fetches by" | tee gpurun_out/r02_s10_e4_default.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "oracle or block_size or plan_of" 2>&1 | tail -3
