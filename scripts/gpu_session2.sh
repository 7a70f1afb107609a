#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== sweep 3 Gbp"; timeout 1500 python tools/sweep.py > gpurun_out/sweep_3g.log 2>&1; echo "rc=$?"; cat gpurun_out/sweep_3g.log
echo "== ncu full 3 Gbp E=0 (auto depth)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:map_kernel -s 3 -c 1 -o gpurun_out/prof_map_3g_e0 -f python tools/sweep.py --configs 0:-1:64 --reps 2 > gpurun_out/ncu_e0.log 2>&1; echo "rc=$?"
echo "== ncu full 3 Gbp E=2"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:map_kernel -s 3 -c 1 -o gpurun_out/prof_map_3g_e2 -f python tools/sweep.py --configs 2:-1:2 --reps 2 > gpurun_out/ncu_e2.log 2>&1; echo "rc=$?"
echo "== bench 3 Gbp"; timeout 1500 python bench.py > gpurun_out/bench_3g.json 2> gpurun_out/bench_3g.log; echo "rc=$?"; tail -4 gpurun_out/bench_3g.log; cat gpurun_out/bench_3g.json
