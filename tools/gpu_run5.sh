#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python - <<'PY' 2>&1 | tee gpurun_out/microbench_solve.log
import sys; sys.path.insert(0, '.')
import geneo4petsc_b200 as g
for h in (256, 1024, 4096, 16384):
    print("solve stream h=%d: %.1f GB/s, %.3f ms" % ((h,) + g.microbench(2, h, 5)))
print("copy GB/s", g.microbench(1, 1 << 28, 10))
PY
GENEO_PROFILE=1 GENEO_PROFILE_OUT=gpurun_out/profile_sites_128.csv python bench.py --size 128 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench128.json 2> gpurun_out/bench128.err; echo "bench128 rc=$?"; cat gpurun_out/bench128.json; tail -5 gpurun_out/bench128.err
python tools/profile_report.py gpurun_out/profile_sites_128.csv | head -14 | tee gpurun_out/profile_128.txt
