#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import geneo4petsc_b200 as g
for n in (1024, 2048, 4096, 8192):
    r = g.microbench(3, n, 5); print("schur-shape gemm n=%d K=128: %.2f TFLOP/s, %.3f ms" % (n, r[0], r[1]))
for n in (4096,):
    r = g.microbench(0, n, 3); print("square dmma gemm n=%d: %.2f TFLOP/s" % (n, r[0]))
PY
GENEO_PROFILE=1 timeout 900 python tools/profile_refactor.py 160 > gpurun_out/profile_refactor_160.log 2>&1; tail -2 gpurun_out/profile_refactor_160.log
python tools/profile_report.py gpurun_out/profile_refactor_160.csv | tee gpurun_out/profile_refactor_160.txt | head -30
