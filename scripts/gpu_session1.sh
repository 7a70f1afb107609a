#!/bin/bash
# First GPU session: parity tests, smoke, random-read microbenchmark, benches, ncu captures.
# Every step has its own timeout and log under gpurun_out/; a failing step does not stop the rest.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
(lscpu | head -25; free -g; df -h . /tmp) > gpurun_out/host.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/smoke.log
echo "== randread"; timeout 300 tools/randread 4096 300 > gpurun_out/randread.log 2>&1; echo "rc=$?"; tail -45 gpurun_out/randread.log
echo "== bench 250 Mbp"; timeout 900 python bench.py --genome-mbp 250 --nchr 5 --seed 44 --steps 4 --warmup 3 --batch-mpos 32 > gpurun_out/bench_250.json 2> gpurun_out/bench_250.log; echo "rc=$?"; tail -5 gpurun_out/bench_250.log; cat gpurun_out/bench_250.json
echo "== ncu launches (250 Mbp)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:map_kernel -c 40 --csv --log-file gpurun_out/launches_250.csv python bench.py --genome-mbp 250 --nchr 5 --seed 44 --steps 2 --warmup 1 --batch-mpos 32 --no-cpu-baseline --extras '' > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full (250 Mbp)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:map_kernel -s 1 -c 1 -o gpurun_out/prof_map_250_e0 -f python bench.py --genome-mbp 250 --nchr 5 --seed 44 --steps 2 --warmup 1 --batch-mpos 32 --no-cpu-baseline --extras '' > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
echo "== bench 3 Gbp"; timeout 1500 python bench.py > gpurun_out/bench_3g.json 2> gpurun_out/bench_3g.log; echo "rc=$?"; tail -8 gpurun_out/bench_3g.log; cat gpurun_out/bench_3g.json
