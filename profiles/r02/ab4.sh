#!/bin/bash
# round 2, GPU call 4 (8 GPUs): multi-rank parity at 2/4/8 ranks, scaling bench at N = 4 and 8 (peer pull), N = 8 on the NCCL
# exchange for comparison, and the 5e8 rows/GPU x 8 = 4e9-record table (configs[4] at 40 % of its named size)
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/ab4_topo.txt 2>&1
SECONDS=0
timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q -s -v > gpurun_out/ab4_tests.log 2>&1; echo "rc=$? after $SECONDS s" >> gpurun_out/ab4_tests.log
grep -h "PASSED\|FAILED\|SKIPPED\|passed\|failed\|rc=" gpurun_out/ab4_tests.log | tail -12
T() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2970$n bench.py --gpus $n "$@"; }
SECONDS=0
timeout 400 bash -c "$(declare -f T); T 8 --steps 10 --warmup 3 --no-cpu" > gpurun_out/ab4_n8.json 2> gpurun_out/ab4_n8.err; echo "N=8 full: $SECONDS s"; tail -c 2600 gpurun_out/ab4_n8.json; tail -3 gpurun_out/ab4_n8.err
SECONDS=0
timeout 300 bash -c "$(declare -f T); T 4 --steps 10 --warmup 3 --no-cpu --no-e2e" > gpurun_out/ab4_n4.json 2> gpurun_out/ab4_n4.err; echo "N=4: $SECONDS s"; tail -c 900 gpurun_out/ab4_n4.json
SECONDS=0
TAD_PEER_PULL=0 timeout 300 bash -c "$(declare -f T); T 8 --steps 10 --warmup 3 --no-cpu --no-e2e" > gpurun_out/ab4_n8_nccl.json 2> gpurun_out/ab4_n8_nccl.err; echo "N=8 nccl: $SECONDS s"; tail -c 900 gpurun_out/ab4_n8_nccl.json
SECONDS=0
timeout 500 bash -c "$(declare -f T); T 8 --series 5000000 --points 100 --steps 3 --warmup 1 --no-cpu --no-e2e" > gpurun_out/ab4_n8_big.json 2> gpurun_out/ab4_n8_big.err; echo "N=8 5e8 rows/GPU: $SECONDS s"; tail -c 1200 gpurun_out/ab4_n8_big.json; tail -3 gpurun_out/ab4_n8_big.err
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/ab4_n*.json")):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print("%-24s value %.3e  %.3f ms" % (p.split("/")[-1], d["value"], d["ms_per_step"]), {k: round(v, 3) for k, v in d["phase_ms"].items() if v}, d.get("parity"))
    except Exception as e:
        print(p, "n/a", e)
PY
