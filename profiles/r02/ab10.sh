#!/bin/bash
# round 2, GPU call 10 (2 GPUs): the full bench line at N = 2 with the sharded side run forced on (the N = 8 line runs it at 5e8 rows per GPU)
set -u
mkdir -p gpurun_out
SECONDS=0
TAD_BENCH_BIG_SIDE=500000 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/ab10_n2_full.json 2> gpurun_out/ab10_n2_full.err
echo "rc=$? after $SECONDS s"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/ab10_n2_full.json").read().strip().splitlines()[-1])
print("value %.3e  %.3f ms  e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["parity"])
for s in d.get("side", []):
    print(json.dumps(s)[:900])
PY
tail -3 gpurun_out/ab10_n2_full.err | cut -c1-300
