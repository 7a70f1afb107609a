#!/bin/bash
# variant entries through the jump tables: parity suite, then walk-only vs variants at 3 Gbp
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/s23_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/s23_pytest.log
for v in 0 1; do
  echo "== sweep GMB_JUMP_VARIANTS=$v"
  GMB_JUMP_VARIANTS=$v timeout 900 python tools/sweep.py --reps 2 --configs 0:-1:256,1:-1:64,2:-1:8,3:-1:0.5 > gpurun_out/s23_sweep_var$v.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/s23_sweep_var$v.log
done
