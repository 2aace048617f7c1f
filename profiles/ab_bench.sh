#!/bin/bash
# Same-box A/B of two builds of the library (run through gpurun, one GPU):
#
#   here:   git stash / checkout the baseline, python -m theia_b200.build --force,
#           cp theia_b200/libtheia_tad.so theia_b200/libtheia_tad_prev.so, go back, rebuild
#   then:   gpurun --timeout 300 -- 'bash profiles/ab_bench.sh'
#
# Runs the GPU tests on the new build, then bench.py (device-resident leg only) on new / prev / new again and prints the
# per-phase milliseconds side by side.  libtheia_tad_prev.so is git-ignored but travels with the snapshot.
set -u
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/ab_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ab_tests.log
tail -3 gpurun_out/ab_tests.log
cp theia_b200/libtheia_tad.so /tmp/ab_new.so
timeout 60 python bench.py --no-cpu --no-e2e > gpurun_out/ab_new.json 2> gpurun_out/ab_new.err
if [ -f theia_b200/libtheia_tad_prev.so ]; then
    cp theia_b200/libtheia_tad_prev.so theia_b200/libtheia_tad.so
    timeout 60 python bench.py --no-cpu --no-e2e > gpurun_out/ab_prev.json 2> gpurun_out/ab_prev.err
    cp /tmp/ab_new.so theia_b200/libtheia_tad.so
    timeout 60 python bench.py --no-cpu --no-e2e > gpurun_out/ab_new2.json 2> gpurun_out/ab_new2.err
fi
python - <<'PY'
import json
for n in ("new", "prev", "new2"):
    try:
        d = json.loads(open("gpurun_out/ab_%s.json" % n).read().strip().splitlines()[-1])
        print("%-5s %.3f ms" % (n, d["ms_per_step"]), {k: round(v, 3) for k, v in d["phase_ms"].items() if v})
    except Exception as e:
        print(n, "n/a", e)
PY
