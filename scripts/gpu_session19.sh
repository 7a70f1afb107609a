#!/bin/bash
# model-chosen part lengths: full parity suite, then equal split vs model at 3 Gbp for E=1,2,3 and K=50, block sizes
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/s19_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/s19_pytest.log
for m in 0 1; do
  echo "== sweep GMB_PART_MODEL=$m"
  GMB_PART_MODEL=$m timeout 900 python tools/sweep.py --reps 2 --configs 0:-1:256,1:-1:64,2:-1:8,2:-1:8:3,2:-1:8:4,2:-1:8:5,2:-1:8:8,3:-1:0.5,3:-1:0.5:3,3:-1:0.5:6 > gpurun_out/s19_sweep_model$m.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/s19_sweep_model$m.log
done
