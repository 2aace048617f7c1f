#!/bin/bash
# round 2, profiling call (1 GPU): launch list + ncu --set full of the shipped kernels (scatter, group x3, direct detector), then the
# default bench line.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
ARGS="bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-parity --no-sides"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python $ARGS > gpurun_out/r02_launches.log 2>&1
tail -2 gpurun_out/r02_launches.log
# one job's worth of the hot kernels: skip the warm-up job's five launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"partition_kernel|group_kernel|detect_ewma" -s 5 -c 5 -f -o gpurun_out/r02_full python $ARGS > gpurun_out/r02_full.log 2>&1
tail -2 gpurun_out/r02_full.log
ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2> /dev/null
ls -la gpurun_out/r02_full.ncu-rep gpurun_out/r02_full_raw.csv
# does bucket-order locality matter on one GPU when the slot span is large (DBSCAN config: 2^19 buckets x 4096-row slots = 64 GB)?
D="timeout 200 python bench.py --algo DBSCAN --series 10000000 --points 24 --steps 3 --warmup 1 --no-cpu --no-e2e"
$D > gpurun_out/r02_dbscan.json 2> gpurun_out/r02_dbscan.err
TAD_SORT_CLASSES=1 $D > gpurun_out/r02_dbscan_sort.json 2> gpurun_out/r02_dbscan_sort.err
python - <<'PY'
import json
for n in ("r02_dbscan", "r02_dbscan_sort"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        print("%-18s %.3f ms" % (n, d["ms_per_step"]), {k: round(v, 3) for k, v in d["phase_ms"].items() if v}, (d.get("parity") or {}).get("ok"))
    except Exception as e:
        print(n, "n/a", e)
PY
SECONDS=0
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_gpu_tests.log 2>&1; echo "rc=$? after $SECONDS s" >> gpurun_out/r02_gpu_tests.log
grep -h "ARIMA vs\|passed\|failed\|rc=" gpurun_out/r02_gpu_tests.log | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
SECONDS=0
timeout 900 python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench_line.err; echo "bench: $SECONDS s"; tail -c 400 gpurun_out/r02_bench_line.json
