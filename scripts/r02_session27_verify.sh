#!/bin/bash
# round 2, session 27: HEAD with the Dna5 N pass verified the way the driver does it — full GPU suite, smoke(), the bench line
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_s27_pytest.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r02_s27_pytest.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_s27_bench_n1.json 2> gpurun_out/r02_s27_bench_n1.log; echo "rc=$?"; python - <<'PY'
import json
j = json.load(open("gpurun_out/r02_s27_bench_n1.json"))
print("E=0 value %.4g e2e %.4g frac %.3f dram_frac %.3f cpu %.4g" % (j["value"], j["e2e"]["value"], j["roofline"]["frac"], j["roofline"]["dram_frac"] or 0, j["cpu_baseline"]["value"]))
for k, v in j["extra"].items(): print(k, "value %.4g e2e %.4g frac %.3f cpu %.4g" % (v["value"], v["e2e"]["value"], v["roofline"]["frac"], v["cpu_baseline"]["value"]))
print({k: v["equal"] for k, v in j["parity"].items()}, j["clocks"])
PY
