#!/bin/bash
# round 2, session 33: the default bench line at HEAD (reads profiles/ncu_traffic.json regenerated in session 32)
mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/r02_s33_bench_n1.json 2> gpurun_out/r02_s33_bench_n1.log; echo "rc=$?"
python - <<'PY'
import json
j = json.load(open("gpurun_out/r02_s33_bench_n1.json"))
print("E=0 value %.4g e2e %.4g frac %.3f dram_frac %.3f cpu %.4g" % (j["value"], j["e2e"]["value"], j["roofline"]["frac"], j["roofline"]["dram_frac"] or 0, j["cpu_baseline"]["value"]), j["roofline"]["traffic_source"])
for k, v in j["extra"].items(): print(k, "value %.4g e2e %.4g frac %.3f cpu %.4g" % (v["value"], v["e2e"]["value"], v["roofline"]["frac"], v["cpu_baseline"]["value"]))
print({k: v["equal"] for k, v in j["parity"].items()}, j["clocks"], "launches", j.get("gpu_launches"))
PY
