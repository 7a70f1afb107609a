#!/bin/bash
# end-of-round evidence: bench line (with cpu_baseline), reference arm, ncu launch list of the bench command,
# DRAM traffic per launch for E=0/1/2, ncu --set full of the E=2 kernel
mkdir -p gpurun_out
nvidia-smi -L
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active"
echo "== bench 3 Gbp"; timeout 1800 python bench.py > gpurun_out/s21_bench_n1.json 2> gpurun_out/s21_bench_n1.log; echo "rc=$?"; tail -4 gpurun_out/s21_bench_n1.log; cat gpurun_out/s21_bench_n1.json
echo "== reference arm"; timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s21_bench_ref.json 2> gpurun_out/s21_bench_ref.log; echo "rc=$?"; cat gpurun_out/s21_bench_ref.json
echo "== ncu launch list of the bench command"; timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:map_kernel -c 60 --csv --log-file gpurun_out/s21_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/s21_ncu_launches.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/s21_launches_bench.csv | cut -c1-300
echo "== ncu traffic, bench E=0 batch"; timeout 1200 ncu --metrics $M --clock-control none -k regex:map_kernel -s 4 -c 1 --csv --log-file gpurun_out/s21_ncu_traffic_e0.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extras '' > /dev/null 2>&1; echo "rc=$?"
echo "== ncu traffic E=1 / E=2 (bench extras batches)"; timeout 900 ncu --metrics $M --clock-control none -k regex:map_kernel -s 3 -c 1 --csv --log-file gpurun_out/s21_ncu_traffic_e1.csv python tools/sweep.py --configs 1:-1:64:0 --reps 2 > /dev/null 2>&1; timeout 900 ncu --metrics $M --clock-control none -k regex:map_kernel -s 3 -c 1 --csv --log-file gpurun_out/s21_ncu_traffic_e2.csv python tools/sweep.py --configs 2:-1:8:0 --reps 2 > /dev/null 2>&1; echo "rc=$?"
echo "== ncu full E=2"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:map_kernel -s 3 -c 1 -o gpurun_out/s21_prof_e2 -f python tools/sweep.py --configs 2:-1:8:0 --reps 2 > gpurun_out/s21_ncu_e2.log 2>&1; echo "rc=$?"
ls -la gpurun_out/ | tail -12
