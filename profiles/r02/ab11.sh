#!/bin/bash
# round 2, GPU call 11 (1 GPU): direct detector with the reciprocal table fetched one sector ahead -- parity tests + A/B against the previous build
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/ab11_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ab11_tests.log
tail -2 gpurun_out/ab11_tests.log
B="timeout 120 python bench.py --no-cpu --no-e2e --no-sides --steps 10 --warmup 3"
$B > gpurun_out/ab11_new.json 2> gpurun_out/ab11_new.err
cp theia_b200/libtheia_tad.so /tmp/new.so; cp theia_b200/libtheia_tad_prev.so theia_b200/libtheia_tad.so
$B > gpurun_out/ab11_prev.json 2> gpurun_out/ab11_prev.err
cp /tmp/new.so theia_b200/libtheia_tad.so
$B > gpurun_out/ab11_new2.json 2> gpurun_out/ab11_new2.err
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/ab11_*.json")):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print("%-18s %.3f ms" % (p.split("/")[-1], d["ms_per_step"]), {k: round(v, 3) for k, v in d["phase_ms"].items() if v}, (d.get("parity") or {}).get("ok"))
    except Exception as e:
        print(p, "n/a", e)
PY
