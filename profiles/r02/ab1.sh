#!/bin/bash
# round 2, GPU call 1: v2 kernels (detect ballot queue, group 5 CTA/SM, scatter L2 knobs) vs round-1 build
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/ab_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ab_tests.log
tail -3 gpurun_out/ab_tests.log
B="timeout 90 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3"
cp theia_b200/libtheia_tad.so /tmp/ab_new.so
$B > gpurun_out/ab_new.json 2> gpurun_out/ab_new.err
cp theia_b200/libtheia_tad_prev.so theia_b200/libtheia_tad.so
$B > gpurun_out/ab_prev.json 2> gpurun_out/ab_prev.err
cp /tmp/ab_new.so theia_b200/libtheia_tad.so
$B > gpurun_out/ab_new2.json 2> gpurun_out/ab_new2.err
for lp in 0 1 2; do for sp in 0 1 2; do
  [ "$lp$sp" = "00" ] && continue
  TAD_SCATTER_LOADPOL=$lp TAD_SCATTER_STOREPOL=$sp $B > gpurun_out/ab_pol_${lp}${sp}.json 2> gpurun_out/ab_pol_${lp}${sp}.err
done; done
TAD_DEBUG_LOGB=10 TAD_OPTIMISTIC=0 timeout 90 python bench.py --no-cpu --no-e2e --steps 3 > gpurun_out/ab_logb10.json 2> gpurun_out/ab_logb10.err
TAD_OPTIMISTIC=0 $B > gpurun_out/ab_exact.json 2> gpurun_out/ab_exact.err
TAD_DETECT_STAGED=0 $B > gpurun_out/ab_unstaged.json 2> gpurun_out/ab_unstaged.err
TAD_DETECT_MODE=1 $B > gpurun_out/ab_direct.json 2> gpurun_out/ab_direct.err
TAD_DETECT_MODE=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/ab_tests_direct.log 2>&1; echo "rc=$?" >> gpurun_out/ab_tests_direct.log
tail -3 gpurun_out/ab_tests_direct.log
for m in 0 1; do
  TAD_ARIMA_MODE=$m timeout 120 python bench.py --algo ARIMA --series 20000 --points 100 --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > gpurun_out/ab_arima_$m.json 2> gpurun_out/ab_arima_$m.err
done
TAD_ARIMA_MODE=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k arima -s > gpurun_out/ab_tests_arima0.log 2>&1; echo "rc=$?" >> gpurun_out/ab_tests_arima0.log
grep -h "ARIMA vs\|passed\|failed" gpurun_out/ab_tests.log gpurun_out/ab_tests_arima0.log | tail -12
(cd profiles/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scatter_tma scatter_tma.cu && timeout 60 ./scatter_tma) > gpurun_out/scatter_tma.log 2>&1
tail -15 gpurun_out/scatter_tma.log
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print("%-28s %.3f ms" % (p.split("/")[-1], d["ms_per_step"]), {k: round(v, 3) for k, v in d["phase_ms"].items() if v})
    except Exception as e:
        print(p, "n/a", e)
PY
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
