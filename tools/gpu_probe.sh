#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/probe_levels.log
import sys; sys.path.insert(0, '.')
import geneo4petsc_b200 as g
# kind = 100*nlev + nr ; n = panel height ; reps = fronts per level
for nr in (1, 8):
    for nf in (1, 8):
        for h in (128, 1024, 8192):
            for nlev in (1, 50):
                gb, ms = g.microbench(100 * nlev + nr, h, nf)
                print("nr=%d fronts/level=%d h=%5d levels=%3d : %8.3f ms  %8.1f GB/s  -> %.1f us per level-phase" % (nr, nf, h, nlev, ms, gb, 1e3 * ms / (2 * nlev)), flush=True)
PY
