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
SECONDS=0
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_gpu_tests.log 2>&1; echo "rc=$? after $SECONDS s" >> gpurun_out/r02_gpu_tests.log
grep -h "ARIMA vs\|passed\|failed\|rc=" gpurun_out/r02_gpu_tests.log | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
SECONDS=0
timeout 900 python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench_line.err; echo "bench: $SECONDS s"; tail -c 400 gpurun_out/r02_bench_line.json
