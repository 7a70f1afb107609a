#!/bin/bash
# round 2, session 16: repeat-rich genomes (30 % of every chromosome covered by planted copies, 2 % / 10 % diverged)
mkdir -p gpurun_out
nvidia-smi -L
echo "== rep-frac 0.3, 2 % diverged"; timeout 900 python tools/sweep.py --genome-mbp 1000 --nchr 8 --rep-frac 0.3 --reps 3 --configs 0:-1:128,1:-1:32,2:-1:4 2>&1 | tee gpurun_out/r02_s16_sweep_rep30.log | grep -v "fetches by"
echo "== the same genome size with the bench's 5 %"; timeout 900 python tools/sweep.py --genome-mbp 1000 --nchr 8 --reps 3 --configs 0:-1:128,1:-1:32,2:-1:4 2>&1 | tee gpurun_out/r02_s16_sweep_rep05.log | grep -v "fetches by"
