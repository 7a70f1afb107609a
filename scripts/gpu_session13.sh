#!/bin/bash
# device run-length encoding for the track writers: parity tests, CLI golden replay, end-to-end writer timing
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest runs + cli"; timeout 1500 python -m pytest tests/test_gpu_cli.py tests/test_gpu_parity.py -m gpu -q -k "runs or cli" > gpurun_out/s13_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/s13_pytest.log
echo "== writers timing 250 Mbp"; timeout 900 bash tools/writers_bench.sh 250 > gpurun_out/s13_writers.log 2>&1; echo "rc=$?"; cat gpurun_out/s13_writers.log
