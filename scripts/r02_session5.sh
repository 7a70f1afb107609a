#!/bin/bash
# round 2, session 5: tuning variants of the new kernels, full GPU suite, bench line (3 Gbp, parity against the reference),
# BASELINE config 5 through bench.py on one GPU
mkdir -p gpurun_out
nvidia-smi -L
echo "== variants"
for v in minb4 batch8; do
  echo "-- $v"; GMB_LIB_PATH=$PWD/genmap_b200/lib/variants/libgenmap_b200_$v.so timeout 600 python tools/sweep.py --reps 3 --configs 1:-1:64,2:-1:8 2>&1 | grep -v "fetches by" | tee gpurun_out/r02_s5_sweep_$v.log
done
echo "-- exact5"; GMB_LIB_PATH=$PWD/genmap_b200/lib/variants/libgenmap_b200_exact5.so timeout 600 python tools/sweep.py --reps 5 --configs 0:-1:256 2>&1 | grep -v "fetches by" | tee gpurun_out/r02_s5_sweep_exact5.log
echo "-- default build"; timeout 600 python tools/sweep.py --reps 3 --configs 0:-1:256,1:-1:64,2:-1:8,3:-1:0.5 2>&1 | grep -v "fetches by" | tee gpurun_out/r02_s5_sweep_default.log
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_s5_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r02_s5_pytest.log
echo "== bench 3 Gbp"; timeout 1500 python bench.py > gpurun_out/r02_s5_bench_n1.json 2> gpurun_out/r02_s5_bench_n1.log; echo "rc=$?"; tail -6 gpurun_out/r02_s5_bench_n1.log; cut -c1-600 gpurun_out/r02_s5_bench_n1.json
echo "== bench config 5 (pan-genome)"; timeout 1500 python bench.py --config pangenome > gpurun_out/r02_s5_bench_pangenome_n1.json 2> gpurun_out/r02_s5_bench_pangenome_n1.log; echo "rc=$?"; tail -6 gpurun_out/r02_s5_bench_pangenome_n1.log; cut -c1-3000 gpurun_out/r02_s5_bench_pangenome_n1.json
