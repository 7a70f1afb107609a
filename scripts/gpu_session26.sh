#!/bin/bash
# end of round: parity suite, A/B of E=0 against the previous library, defaults sweeps, bench line + reference arm,
# ncu launch list and DRAM traffic of the bench kernels
mkdir -p gpurun_out
nvidia-smi -L
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active"
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/s26_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/s26_pytest.log
echo "== previous library E=0"; GMB_LIB_PATH=$PWD/genmap_b200/lib/variants/libgenmap_b200_prev.so timeout 600 python tools/sweep.py --reps 3 --configs 0:-1:256 > gpurun_out/s26_sweep_prev.log 2>&1; grep -v "fetches by" gpurun_out/s26_sweep_prev.log
echo "== current library"; timeout 900 python tools/sweep.py --reps 3 --configs 0:-1:256,1:-1:64,2:-1:8,3:-1:0.5 > gpurun_out/s26_sweep.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/s26_sweep.log
echo "== dna5"; timeout 900 python tools/sweep.py --n-frac 0.05 --reps 2 --configs 0:-1:256,1:-1:64,2:-1:8 > gpurun_out/s26_sweep_dna5.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/s26_sweep_dna5.log
echo "== bench 3 Gbp"; timeout 1800 python bench.py > gpurun_out/s26_bench_n1.json 2> gpurun_out/s26_bench_n1.log; echo "rc=$?"; tail -3 gpurun_out/s26_bench_n1.log; cat gpurun_out/s26_bench_n1.json | cut -c1-1200
echo "== ncu launch list"; timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:map_kernel -c 60 --csv --log-file gpurun_out/s26_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/s26_ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu traffic E=1 / E=2"; timeout 900 ncu --metrics $M --clock-control none -k regex:map_kernel -s 3 -c 1 --csv --log-file gpurun_out/s26_ncu_traffic_e1.csv python tools/sweep.py --configs 1:-1:64:0 --reps 2 > /dev/null 2>&1; timeout 900 ncu --metrics $M --clock-control none -k regex:map_kernel -s 3 -c 1 --csv --log-file gpurun_out/s26_ncu_traffic_e2.csv python tools/sweep.py --configs 2:-1:8:0 --reps 2 > /dev/null 2>&1; echo "rc=$?"
