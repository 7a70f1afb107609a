#!/bin/bash
# search-scheme part lengths at 3 Gbp: E=2 and E=1 (and K=50 E=2)
mkdir -p gpurun_out
nvidia-smi -L
echo "== E=2"; timeout 900 python tools/parts_sweep.py -E 2 --blocks 6,4 > gpurun_out/s18_parts_e2.log 2>&1; echo "rc=$?"; cat gpurun_out/s18_parts_e2.log
echo "== E=1"; timeout 600 python tools/parts_sweep.py -E 1 --blocks 4 --batch-mpos 64 --weights "1,1;6,7;7,6;5,7;7,5;4,5;5,4" > gpurun_out/s18_parts_e1.log 2>&1; echo "rc=$?"; cat gpurun_out/s18_parts_e1.log
