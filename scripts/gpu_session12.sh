#!/bin/bash
# where do the rank-block fetches go?  histogram by interval size + thin-path statistics at 3 Gbp
mkdir -p gpurun_out
nvidia-smi -L
echo "== sweep hist"; timeout 900 python tools/sweep.py --reps 2 --configs 0:-1:256,1:-1:64,1:-1:64:1,2:-1:8,2:-1:8:1,2:-1:8:3 > gpurun_out/s12_hist.log 2>&1; echo "rc=$?"; cat gpurun_out/s12_hist.log
