#!/bin/bash
# round 2, GPU call 9 (4 GPUs): new defaults at N = 4 (peer pull + sorted class lists), with and without the larger buckets; multi-rank tests at 2 / 4 ranks
set -u
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29704 bench.py --gpus 4 --steps 6 --warmup 2 --no-cpu --no-e2e"
timeout 200 $T > gpurun_out/ab9_n4_default.json 2> gpurun_out/ab9_n4_default.err
TAD_GROUP_TARGET=1536 timeout 200 $T > gpurun_out/ab9_n4_big.json 2> gpurun_out/ab9_n4_big.err
timeout 400 python -m pytest tests/test_multi_rank.py -m gpu -x -q > gpurun_out/ab9_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ab9_tests.log
tail -3 gpurun_out/ab9_tests.log
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/ab9_*.json")):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print("%-26s value %.3e  %.3f ms" % (p.split("/")[-1], d["value"], d["ms_per_step"]), {k: round(v, 3) for k, v in d["phase_ms"].items() if v}, (d.get("parity") or {}).get("ok"))
    except Exception as e:
        print(p, "n/a", e)
PY
