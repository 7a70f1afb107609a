#!/bin/bash
mkdir -p gpurun_out
CFG="0:-1:256,0:15:256,1:-1:64,2:-1:8"
echo "== sweep mb4 (default build)"; timeout 900 python tools/sweep.py --configs $CFG > gpurun_out/sweep_mb4.log 2>&1; echo "rc=$?"; cat gpurun_out/sweep_mb4.log
echo "== sweep mb4 L2 fetch 32"; GMB_L2_FETCH=32 timeout 900 python tools/sweep.py --configs $CFG > gpurun_out/sweep_mb4_l2f32.log 2>&1; echo "rc=$?"; cat gpurun_out/sweep_mb4_l2f32.log
echo "== sweep mb4 L2 fetch 128"; GMB_L2_FETCH=128 timeout 900 python tools/sweep.py --configs $CFG > gpurun_out/sweep_mb4_l2f128.log 2>&1; echo "rc=$?"; cat gpurun_out/sweep_mb4_l2f128.log
echo "== sweep mb3"; GMB_LIB_PATH=$PWD/build/variants/libgmb_mb3.so timeout 900 python tools/sweep.py --configs $CFG > gpurun_out/sweep_mb3.log 2>&1; echo "rc=$?"; cat gpurun_out/sweep_mb3.log
echo "== sweep mb5"; GMB_LIB_PATH=$PWD/build/variants/libgmb_mb5.so timeout 900 python tools/sweep.py --configs $CFG > gpurun_out/sweep_mb5.log 2>&1; echo "rc=$?"; cat gpurun_out/sweep_mb5.log
echo "== bench 3 Gbp"; timeout 1500 python bench.py > gpurun_out/bench_3g.json 2> gpurun_out/bench_3g.log; echo "rc=$?"; tail -4 gpurun_out/bench_3g.log; cat gpurun_out/bench_3g.json
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu.log
