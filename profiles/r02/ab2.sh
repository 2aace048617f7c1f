#!/bin/bash
# round 2, GPU call 2 (1 GPU): r01 kernels restored (v2 detect/scatter were slower: profiles/r02/ab1_summary.txt), direct detector,
# evaluation-synchronous ARIMA, 4-worker C host under compute-sanitizer, full bench line
set -u
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q -s > gpurun_out/ab2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ab2_tests.log
grep -h "ARIMA vs\|passed\|failed\|rc=" gpurun_out/ab2_tests.log | tail -8
B="timeout 120 python bench.py --no-cpu --no-e2e --no-parity --steps 10 --warmup 3"
$B > gpurun_out/ab2_staged.json 2> gpurun_out/ab2_staged.err
TAD_DETECT_MODE=1 $B > gpurun_out/ab2_direct.json 2> gpurun_out/ab2_direct.err
$B > gpurun_out/ab2_staged2.json 2> gpurun_out/ab2_staged2.err
TAD_DETECT_MODE=1 $B > gpurun_out/ab2_direct2.json 2> gpurun_out/ab2_direct2.err
TAD_DETECT_MODE=1 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_host_mirror.py -m gpu -x -q > gpurun_out/ab2_tests_direct.log 2>&1; echo "rc=$?" >> gpurun_out/ab2_tests_direct.log
tail -3 gpurun_out/ab2_tests_direct.log
for m in 0 1; do
  TAD_ARIMA_MODE=$m timeout 200 python bench.py --algo ARIMA --series 20000 --points 100 --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > gpurun_out/ab2_arima_$m.json 2> gpurun_out/ab2_arima_$m.err
done
TAD_ARIMA_MODE=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k arima -s > gpurun_out/ab2_tests_arima0.log 2>&1; echo "rc=$?" >> gpurun_out/ab2_tests_arima0.log
grep -h "ARIMA vs\|passed\|failed" gpurun_out/ab2_tests_arima0.log | tail -6
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/ab2_*.json")):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print("%-28s %.3f ms" % (p.split("/")[-1], d["ms_per_step"]), {k: round(v, 3) for k, v in d["phase_ms"].items() if v})
    except Exception as e:
        print(p, "n/a", e)
PY
# 4 workers on one context under compute-sanitizer (memcheck)
gcc -std=c99 -pthread -Iinclude examples/tad_workers.c -Ltheia_b200 -ltheia_tad -Wl,-rpath,$PWD/theia_b200 -o /tmp/tad_workers
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 /tmp/tad_workers > gpurun_out/ab2_sanitizer_workers.log 2>&1; echo "rc=$?" >> gpurun_out/ab2_sanitizer_workers.log
tail -6 gpurun_out/ab2_sanitizer_workers.log
(cd profiles/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scatter_tma scatter_tma.cu && timeout 60 ./scatter_tma) > gpurun_out/ab2_scatter_tma.log 2>&1
tail -12 gpurun_out/ab2_scatter_tma.log
# the full default bench line (sides, parity, cpu baseline) and the reference arm, timed
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/ab2_bench_full.json 2> gpurun_out/ab2_bench_full.err; tail -c 1500 gpurun_out/ab2_bench_full.json
grep -h "Elapsed\|Maximum resident" gpurun_out/ab2_bench_full.err
/usr/bin/time -v timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ab2_bench_ref.json 2> gpurun_out/ab2_bench_ref.err; tail -c 600 gpurun_out/ab2_bench_ref.json
grep -h "Elapsed" gpurun_out/ab2_bench_ref.err
