#!/bin/bash
mkdir -p gpurun_out
echo "== ldflavor rates"; timeout 300 tools/ldflavor 3072 200 > gpurun_out/ldflavor_rates.log 2>&1; cat gpurun_out/ldflavor_rates.log
echo "== ldflavor traffic (ncu)"; timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum --clock-control none -k regex:chase --csv --log-file gpurun_out/ldflavor_ncu.csv tools/ldflavor 3072 100 > /dev/null 2>&1; echo "rc=$?"
python3 - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/ldflavor_ncu.csv')))
hdr=None; out={}
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr is None or len(r)<len(hdr): continue
    d=dict(zip(hdr,r))
    out.setdefault((d['ID'],d['Kernel Name']),{})[d['Metric Name']]=d['Metric Value']
for (i,k),m in out.items(): print(i,k[:40],m)
PY
