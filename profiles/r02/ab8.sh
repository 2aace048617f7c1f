#!/bin/bash
# round 2, GPU call 8 (8 GPUs): what limits the peer pull at N = 8?  Sorted class lists (page locality), larger buckets (fewer, larger
# pulled pieces), the dense exact partition pulled by the peers.
set -u
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29708 bench.py --gpus 8 --steps 6 --warmup 2 --no-cpu --no-e2e"
TAD_PEER_PULL=1 TAD_SORT_CLASSES=1 timeout 200 $T > gpurun_out/ab8_pull_sort.json 2> gpurun_out/ab8_pull_sort.err
TAD_PEER_PULL=1 TAD_SORT_CLASSES=1 TAD_GROUP_TARGET=1536 timeout 200 $T > gpurun_out/ab8_pull_sort_big.json 2> gpurun_out/ab8_pull_sort_big.err
TAD_PEER_PULL=1 TAD_OPTIMISTIC=0 TAD_EXACT_PULL=1 TAD_SORT_CLASSES=1 timeout 200 $T > gpurun_out/ab8_xpull_sort.json 2> gpurun_out/ab8_xpull_sort.err
TAD_PEER_PULL=1 TAD_OPTIMISTIC=0 TAD_EXACT_PULL=1 timeout 200 $T > gpurun_out/ab8_xpull.json 2> gpurun_out/ab8_xpull.err
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/ab8_*.json")):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print("%-26s value %.3e  %.3f ms" % (p.split("/")[-1], d["value"], d["ms_per_step"]), {k: round(v, 3) for k, v in d["phase_ms"].items() if v}, (d.get("parity") or {}).get("ok"))
    except Exception as e:
        print(p, "n/a", e)
PY
tail -3 gpurun_out/ab8_*.err | cut -c1-300 | tail -20
