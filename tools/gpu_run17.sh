#!/bin/bash
mkdir -p gpurun_out
for ST in 3 2; do
echo "=== GENEO_GEMM_STAGES=$ST"
GENEO_GEMM_STAGES=$ST python - <<'PY'
import geneo4petsc_b200 as g
for n in (200, 513):
    r = g.microbench(0, n, 1); print("dmma gemm n=%d err %.2e" % (n, r[1]))
for n in (2048, 8192):
    r = g.microbench(3, n, 5); print("schur-shape gemm n=%d K=128: %.2f TFLOP/s, %.3f ms" % (n, r[0], r[1]))
r = g.microbench(0, 4096, 3); print("square dmma gemm n=4096: %.2f TFLOP/s" % r[0])
PY
GENEO_GEMM_STAGES=$ST GENEO_HOSTPROF=1 timeout 800 python tools/profile_refactor.py 160 2>&1 | grep "factorize: host side total\|refactor " | tail -2 | cut -c1-120
done
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
