#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
CFG="0:-1:256:1,1:-1:64:1,1:-1:64:4,1:-1:64:6,1:-1:64:8,1:-1:64:12,2:-1:16:1,2:-1:16:3,2:-1:16:4,2:-1:16:6,2:-1:16:8,2:-1:16:10"
echo "== sweep block sizes"; timeout 1500 python tools/sweep.py --configs $CFG > gpurun_out/sweep_blocks.log 2>&1; echo "rc=$?"; cat gpurun_out/sweep_blocks.log
