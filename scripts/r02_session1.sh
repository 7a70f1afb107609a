#!/bin/bash
# round 2, session 1: parity at the benchmarked scale (deep jump tables, 3 Gbp plan on small genomes, 250 Mbp cmp),
# bench line with `parity`, ncu counters of the three bench kernels at HEAD
mkdir -p gpurun_out
nvidia-smi -L
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum"
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r02_s1_pytest.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r02_s1_pytest.log
echo "== bench 3 Gbp"; timeout 1500 python bench.py > gpurun_out/r02_s1_bench_n1.json 2> gpurun_out/r02_s1_bench_n1.log; echo "rc=$?"; tail -5 gpurun_out/r02_s1_bench_n1.log; cut -c1-1500 gpurun_out/r02_s1_bench_n1.json
echo "== ncu counters E=0/1/2"
for cfg in 0:-1:256 1:-1:64 2:-1:8; do
  E=${cfg%%:*}
  timeout 600 ncu --metrics $M --clock-control none -k regex:map_kernel -s 3 -c 1 --csv --log-file gpurun_out/r02_s1_ncu_e$E.csv python tools/sweep.py --configs $cfg --reps 2 > gpurun_out/r02_s1_ncu_e$E.log 2>&1; echo "E=$E rc=$?"
done
