#!/bin/bash
# Dna5 tuning: resident CTAs per SM (register budget) x non-blocked instantiation, parity first
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest dna5"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "dna5 or golden or locations" > gpurun_out/s15_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/s15_pytest.log
for v in "" mb5_3 mb5_4; do
  echo "== sweep dna5 variant '$v'"
  if [ -n "$v" ]; then export GMB_LIB_PATH=$PWD/genmap_b200/lib/variants/libgenmap_b200_$v.so; fi
  timeout 600 python tools/sweep.py --n-frac 0.05 --reps 2 --configs 0:-1:256,1:-1:64,2:-1:8 > gpurun_out/s15_sweep_dna5_$v.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/s15_sweep_dna5_$v.log
done
