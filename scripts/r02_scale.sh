#!/bin/bash
# bench.py at N GPUs, launched the way the driver launches it:  bash scripts/r02_scale.sh N
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r02_scale_n$N.out 2> gpurun_out/r02_scale_n$N.log
echo "rc=$?"; grep "^{" gpurun_out/r02_scale_n$N.out > gpurun_out/r02_scale_n$N.json
python - <<PY
import json
j = json.load(open("gpurun_out/r02_scale_n$N.json"))
print("N=$N E=0 value %.4g e2e %.4g d2h-only %.4g" % (j["value"], j["e2e"]["value"], j["e2e"]["d2h_only_positions_per_s"]))
for k, v in j["extra"].items(): print("   ", k, "value %.4g e2e %.4g" % (v["value"], v["e2e"]["value"]))
print("   ", j["config"]["sharding"])
PY
