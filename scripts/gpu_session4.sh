#!/bin/bash
mkdir -p gpurun_out
CFG="0:-1:256,0:15:256,1:-1:64,2:-1:8"
echo "== sweep W=1 mb4 (default build)"; timeout 900 python tools/sweep.py --configs $CFG > gpurun_out/sweep_w1_mb4.log 2>&1; echo "rc=$?"; cat gpurun_out/sweep_w1_mb4.log
for v in w1_mb5 w1_mb3 w3_mb4; do
  echo "== sweep $v"; GMB_LIB_PATH=$PWD/build/variants/libgmb_$v.so timeout 900 python tools/sweep.py --configs $CFG > gpurun_out/sweep_$v.log 2>&1; echo "rc=$?"; cat gpurun_out/sweep_$v.log
done
echo "== ncu full 3 Gbp E=2 (W=1)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:map_kernel -s 3 -c 1 -o gpurun_out/prof_w1_3g_e2 -f python tools/sweep.py --configs 2:-1:2 --reps 2 > gpurun_out/ncu_e2.log 2>&1; echo "rc=$?"
echo "== ncu full 3 Gbp E=0 (W=1)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:map_kernel -s 3 -c 1 -o gpurun_out/prof_w1_3g_e0 -f python tools/sweep.py --configs 0:-1:64 --reps 2 > gpurun_out/ncu_e0.log 2>&1; echo "rc=$?"
echo "== bench 3 Gbp"; timeout 1800 python bench.py > gpurun_out/bench_3g.json 2> gpurun_out/bench_3g.log; echo "rc=$?"; tail -6 gpurun_out/bench_3g.log; cat gpurun_out/bench_3g.json
echo "== bench --impl reference"; timeout 1800 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.log; echo "rc=$?"; tail -6 gpurun_out/bench_ref.log; cat gpurun_out/bench_ref.json
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu.log
