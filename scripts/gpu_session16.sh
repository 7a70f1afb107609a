#!/bin/bash
# -ep beyond 64 files; full suite after the locate refactor and the Dna5 launch bounds; Dna5 sweep with the final defaults
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/s16_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/s16_pytest.log
echo "== sweep dna5 final"; timeout 600 python tools/sweep.py --n-frac 0.05 --reps 2 --configs 0:-1:256,1:-1:64,2:-1:8 > gpurun_out/s16_sweep_dna5.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/s16_sweep_dna5.log
