#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
python - <<'PY'
import geneo4petsc_b200 as g
for nr in (8, 16):
    r = g.microbench(100 * 1 + nr, 4096, 400); print("nr=%d single level 400 fronts h=4096: %8.3f ms %8.1f GB/s" % (nr, r[1], r[0]))
    r = g.microbench(100 * 50 + nr, 8192, 1); print("nr=%d 1 front/level h=8192 x50 levels: %8.3f ms -> %.1f us per level-phase" % (nr, r[1], r[1] * 10))
PY
for B in 8 16; do
GENEO_EXTRA_OPTS="-els2_eps_block $B" GENEO_HOSTPROF=1 timeout 800 python tools/profile_refactor.py 160 2>&1 | grep "lanczos: all\|refactor \|launch_solve\|eig steps" | tail -4 | cut -c1-150
done
