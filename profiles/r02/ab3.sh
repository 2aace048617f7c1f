#!/bin/bash
# round 2, GPU call 3 (2 GPUs): peer-pull multi-GPU path -- parity tests, then bench at N = 2 (default, and with TAD_PEER_PULL=0)
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/ab3_topo.txt 2>&1
timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -x -q -s > gpurun_out/ab3_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ab3_tests.log
tail -25 gpurun_out/ab3_tests.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2"
timeout 300 $T --steps 10 --warmup 3 --no-cpu > gpurun_out/ab3_n2.json 2> gpurun_out/ab3_n2.err; tail -c 2500 gpurun_out/ab3_n2.json; tail -5 gpurun_out/ab3_n2.err
TAD_PEER_PULL=0 timeout 300 $T --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab3_n2_nccl.json 2> gpurun_out/ab3_n2_nccl.err; tail -c 800 gpurun_out/ab3_n2_nccl.json
TAD_DETECT_MODE=1 timeout 300 $T --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab3_n2_direct.json 2> gpurun_out/ab3_n2_direct.err; tail -c 800 gpurun_out/ab3_n2_direct.json
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab3_n1.json 2> gpurun_out/ab3_n1.err; tail -c 600 gpurun_out/ab3_n1.json
# ---- single-GPU experiments riding the same call -----------------------------------------------------------------
B="timeout 120 python bench.py --no-cpu --no-e2e --no-parity --steps 10 --warmup 3"
TAD_SCATTER_RPT=4 $B > gpurun_out/ab3_rpt4.json 2> gpurun_out/ab3_rpt4.err
$B > gpurun_out/ab3_rpt8.json 2> gpurun_out/ab3_rpt8.err
TAD_SCATTER_RPT=4 $B > gpurun_out/ab3_rpt4b.json 2> gpurun_out/ab3_rpt4b.err
TAD_SCATTER_RPT=4 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/ab3_tests_rpt4.log 2>&1; echo "rc=$?" >> gpurun_out/ab3_tests_rpt4.log
tail -3 gpurun_out/ab3_tests_rpt4.log
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/ab3_*.json")):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print("%-28s %.3f ms" % (p.split("/")[-1], d["ms_per_step"]), {k: round(v, 3) for k, v in d["phase_ms"].items() if v}, d.get("parity"))
    except Exception as e:
        print(p, "n/a", e)
PY
SECONDS=0
timeout 900 python bench.py > gpurun_out/ab3_bench_full.json 2> gpurun_out/ab3_bench_full.err; echo "full bench: $SECONDS s"; tail -c 3000 gpurun_out/ab3_bench_full.json; tail -3 gpurun_out/ab3_bench_full.err
SECONDS=0
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ab3_bench_ref.json 2> gpurun_out/ab3_bench_ref.err; echo "reference arm: $SECONDS s"; tail -c 900 gpurun_out/ab3_bench_ref.json
