#!/bin/bash
# round 2, final evidence session (1 GPU): full GPU suite, smoke, bench line + reference arm, ncu launch list of the
# bench command, DRAM traffic per launch of the three bench kernels (-> profiles/ncu_traffic.json), --set full of the
# headline kernel
mkdir -p gpurun_out
SHA=$1
nvidia-smi -L
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum"
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_s12_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r02_s12_pytest.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== ncu traffic of the bench kernels (second timed launch of each configuration)"
timeout 600 ncu --metrics $M --clock-control none -k regex:exact_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02_s12_ncu_e0.csv python tools/sweep.py --configs 0:-1:256 --reps 2 > gpurun_out/r02_s12_ncu_e0.log 2>&1; echo "E=0 rc=$?"
timeout 600 ncu --metrics $M --clock-control none -k regex:block_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02_s12_ncu_e1.csv python tools/sweep.py --configs 1:-1:64 --reps 2 > gpurun_out/r02_s12_ncu_e1.log 2>&1; echo "E=1 rc=$?"
timeout 600 ncu --metrics $M --clock-control none -k regex:block_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02_s12_ncu_e2.csv python tools/sweep.py --configs 2:-1:8 --reps 2 > gpurun_out/r02_s12_ncu_e2.log 2>&1; echo "E=2 rc=$?"
grep -h "^E=" gpurun_out/r02_s12_ncu_e?.log
python scripts/update_ncu_traffic.py $SHA K30_E0_batch268435456_genome3000000000=gpurun_out/r02_s12_ncu_e0.csv K30_E1_batch67108864_genome3000000000=gpurun_out/r02_s12_ncu_e1.csv K30_E2_batch8388608_genome3000000000=gpurun_out/r02_s12_ncu_e2.csv > /dev/null && cp profiles/ncu_traffic.json gpurun_out/r02_s12_ncu_traffic.json
echo "== bench 3 Gbp (reads the traffic just measured)"; timeout 1500 python bench.py > gpurun_out/r02_s12_bench_n1.json 2> gpurun_out/r02_s12_bench_n1.log; echo "rc=$?"; tail -4 gpurun_out/r02_s12_bench_n1.log; cut -c1-300 gpurun_out/r02_s12_bench_n1.json
echo "== reference arm"; timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_s12_bench_ref.json 2> gpurun_out/r02_s12_bench_ref.log; echo "rc=$?"; cut -c1-400 gpurun_out/r02_s12_bench_ref.json
echo "== ncu launch list of the bench command"; timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:exact_kernel|block_kernel|map_kernel|k_jump|k_locate" -c 80 --csv --log-file gpurun_out/r02_s12_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02_s12_ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu --set full of the headline kernel"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:exact_kernel -s 1 -c 1 -o gpurun_out/r02_s12_full_e0 python tools/sweep.py --configs 0:-1:256 --reps 2 > gpurun_out/r02_s12_ncu_full.log 2>&1; echo "rc=$?"; ls -la gpurun_out/r02_s12_full_e0.ncu-rep
timeout 300 ncu -i gpurun_out/r02_s12_full_e0.ncu-rep --page details > gpurun_out/r02_s12_full_e0_details.txt 2>&1; timeout 300 ncu -i gpurun_out/r02_s12_full_e0.ncu-rep --page raw --csv > gpurun_out/r02_s12_full_e0_raw.csv 2>&1; wc -l gpurun_out/r02_s12_full_e0_details.txt
