#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for G in "" 1; do
echo "=== GENEO_SOLVE_GENERIC=$G"
GENEO_SOLVE_GENERIC=$G timeout 300 python - <<'PY'
import geneo4petsc_b200 as g
for nr in (1, 8):
  for nf in (1, 8):
    for h in (1024, 8192):
        r = g.microbench(100 * 50 + nr, h, nf); print("nr=%d fronts/level=%d h=%5d levels=50 : %8.3f ms %8.1f GB/s -> %.1f us per level-phase" % (nr, nf, h, r[1], r[0], r[1] * 1e3 / 100))
  r = g.microbench(100 * 1 + nr, 4096, 400); print("nr=%d single level 400 fronts h=4096: %8.3f ms %8.1f GB/s" % (nr, r[1], r[0]))
PY
done > gpurun_out/probe_nr8.log 2>&1
cat gpurun_out/probe_nr8.log
timeout 900 python bench.py --size 128 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench128.json 2> gpurun_out/bench128.err; echo "bench128 rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench128.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['achieved'], d['detail']['numeric_phases_s'], d['detail']['iter_s'], d['detail']['iterations'], d['detail']['dimE'])
PY
tail -3 gpurun_out/bench128.err
