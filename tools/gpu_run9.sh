#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python - > gpurun_out/probe_ring.log 2>&1 <<'PY'
import geneo4petsc_b200 as g
for h in (1024, 16384):
    r = g.microbench(2, h, 5); print("solve stream h=%d: %.1f GB/s, %.3f ms" % (h, r[0], r[1]))
for nf in (1, 8):
    for h in (128, 1024, 8192):
        r = g.microbench(100 * 50 + 1, h, nf); print("nr=1 fronts/level=%d h=%5d levels=50 : %8.3f ms %8.1f GB/s -> %.1f us per level-phase" % (nf, h, r[1], r[0], r[1] * 1e3 / 100))
PY
cat gpurun_out/probe_ring.log
timeout 600 python tools/level_profile.py 128 8 > gpurun_out/levels_128.log 2>&1; head -3 gpurun_out/levels_128.log; tail -3 gpurun_out/levels_128.log
if [ -n "$DO200" ]; then timeout 900 python tools/level_profile.py 200 8 > gpurun_out/levels_200.log 2>&1; head -3 gpurun_out/levels_200.log; tail -3 gpurun_out/levels_200.log; fi
