#!/bin/bash
# round 2, GPU call 7 (2 GPUs): sorted class lists (TAD_SORT_CLASSES=1) at N = 1 and N = 2
set -u
mkdir -p gpurun_out
B="timeout 120 python bench.py --no-cpu --no-e2e --no-sides --steps 10 --warmup 3"
TAD_SORT_CLASSES=1 $B > gpurun_out/ab7_n1_sort.json 2> gpurun_out/ab7_n1_sort.err
$B > gpurun_out/ab7_n1.json 2> gpurun_out/ab7_n1.err
TAD_SORT_CLASSES=1 TAD_GROUP_CONCURRENT=1 $B > gpurun_out/ab7_n1_sort_conc.json 2> gpurun_out/ab7_n1_sort_conc.err
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-e2e"
TAD_SORT_CLASSES=1 timeout 300 $T > gpurun_out/ab7_n2_sort.json 2> gpurun_out/ab7_n2_sort.err
TAD_SORT_CLASSES=1 TAD_OPTIMISTIC=0 TAD_EXACT_PULL=1 timeout 300 $T > gpurun_out/ab7_n2_xpull_sort.json 2> gpurun_out/ab7_n2_xpull_sort.err
TAD_SORT_CLASSES=1 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_multi_rank.py -m gpu -x -q > gpurun_out/ab7_tests_sort.log 2>&1; echo "rc=$?" >> gpurun_out/ab7_tests_sort.log
tail -3 gpurun_out/ab7_tests_sort.log
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/ab7_*.json")):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print("%-28s value %.3e  %.3f ms" % (p.split("/")[-1], d["value"], d["ms_per_step"]), {k: round(v, 3) for k, v in d["phase_ms"].items() if v}, (d.get("parity") or {}).get("ok"))
    except Exception as e:
        print(p, "n/a", e)
PY
