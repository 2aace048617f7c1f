#!/bin/bash
# round 2, GPU call 5 (2 GPUs): N = 2 again after the OMP_PROC_BIND leak was removed from the GPU process (call 3: every rank's host
# threads were pinned to one core -> 7.6 ms of barrier skew, NCCL exchange 6x slower than in round 1)
set -u
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2"
timeout 300 $T --steps 10 --warmup 3 --no-cpu > gpurun_out/ab5_n2.json 2> gpurun_out/ab5_n2.err; tail -3 gpurun_out/ab5_n2.err
TAD_PEER_PULL=0 timeout 300 $T --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab5_n2_nccl.json 2> gpurun_out/ab5_n2_nccl.err
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-sides > gpurun_out/ab5_n1.json 2> gpurun_out/ab5_n1.err
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/ab5_*.json")):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print("%-24s value %.3e  %.3f ms" % (p.split("/")[-1], d["value"], d["ms_per_step"]), {k: round(v, 3) for k, v in d["phase_ms"].items() if v}, d.get("parity"), d.get("e2e"), d.get("config", {}).get("timing"))
    except Exception as e:
        print(p, "n/a", e)
PY
