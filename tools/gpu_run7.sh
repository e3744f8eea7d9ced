#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python - <<'PY' 2>&1 | tee gpurun_out/probe_levels2.log
import sys; sys.path.insert(0, '.')
import geneo4petsc_b200 as g
for h in (1024, 16384):
    print("solve stream h=%d: %.1f GB/s, %.3f ms" % ((h,) + g.microbench(2, h, 5)))
for nr in (1,):
    for nf in (1, 8):
        for h in (128, 1024, 8192):
            for nlev in (50,):
                gb, ms = g.microbench(100 * nlev + nr, h, nf)
                print("nr=%d fronts/level=%d h=%5d levels=%3d : %8.3f ms  %8.1f GB/s  -> %.1f us per level-phase" % (nr, nf, h, nlev, ms, gb, 1e3 * ms / (2 * nlev)), flush=True)
PY
python bench.py --size 128 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench128.json 2> gpurun_out/bench128.err; echo "bench128 rc=$?"; cat gpurun_out/bench128.json; tail -5 gpurun_out/bench128.err
