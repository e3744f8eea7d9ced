#!/bin/bash
mkdir -p gpurun_out
for F in 0 1 2 4 3 7; do
echo "=== GENEO_RING_FLAGS=$F"
GENEO_RING_FLAGS=$F timeout 300 python - <<'PY'
import geneo4petsc_b200 as g
for h in (1024, 16384):
    r = g.microbench(2, h, 5); print("solve stream h=%d: %.1f GB/s, %.3f ms" % (h, r[0], r[1]))
for nf in (1, 8):
    for h in (1024, 8192):
        r = g.microbench(100 * 50 + 1, h, nf); print("nr=1 fronts/level=%d h=%5d levels=50 : %8.3f ms %8.1f GB/s -> %.1f us per level-phase" % (nf, h, r[1], r[0], r[1] * 1e3 / 100))
PY
done > gpurun_out/ring_flags.log 2>&1
cat gpurun_out/ring_flags.log
for F in 1 2; do
GENEO_RING_FLAGS=$F timeout 600 python tools/level_profile.py 128 8 > gpurun_out/levels_128_f$F.log 2>&1; echo "flags $F"; head -2 gpurun_out/levels_128_f$F.log; tail -2 gpurun_out/levels_128_f$F.log
done
