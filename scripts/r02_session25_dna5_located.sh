#!/bin/bash
# round 2, session 25: why the Dna5 sweep of session 24 showed no located entries at E >= 1 — device counters against the
# host mirror on a small genome (new test), then the 3 Gbp Dna5 genome at E = 1 with the table depth and block size forced
mkdir -p gpurun_out
nvidia-smi -L
echo "== located entries: device vs host mirror"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "located_entries_are_used or fetch_counter" 2>&1 | tail -15
echo "== Dna5 (5 % N) at E = 1, E = 2: depth / block size forced"; timeout 600 python tools/sweep.py --n-frac 0.05 --reps 2 --configs 1:-1:64,1:-1:64:1,1:16:64:1,1:14:64:3,1:12:64:3,2:-1:8,2:12:8:4 2>&1 | tee gpurun_out/r02_s25_sweep_dna5_depths.log
