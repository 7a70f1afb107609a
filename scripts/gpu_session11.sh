#!/bin/bash
# csv / locations on the device: full GPU parity suite (no -x, so every failure shows), then a first timing of the locate path
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/s11_pytest.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/s11_pytest.log
echo "== locate timing"; timeout 600 python tools/locate_bench.py > gpurun_out/s11_locate.log 2>&1; echo "rc=$?"; cat gpurun_out/s11_locate.log
