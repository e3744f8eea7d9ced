#!/bin/bash
mkdir -p gpurun_out
probe() {
timeout 300 python - <<'PY'
import geneo4petsc_b200 as g
for nr in (8,):
  for nf in (1, 8):
    for h in (1024, 8192):
        r = g.microbench(100 * 50 + nr, h, nf); print("nr=%d fronts/level=%d h=%5d levels=50 : %8.3f ms %8.1f GB/s -> %.1f us per level-phase" % (nr, nf, h, r[1], r[0], r[1] * 1e3 / 100))
  r = g.microbench(100 * 1 + nr, 4096, 400); print("nr=%d single level 400 fronts h=4096: %8.3f ms %8.1f GB/s" % (nr, r[1], r[0]))
PY
}
echo "=== ring"; probe
echo "=== generic"; GENEO_SOLVE_GENERIC=1 probe
summ() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['achieved'], d['detail']['numeric_phases_s'], d['detail']['iter_s'], d['detail']['iterations'], d['detail']['dimE'])
PY
}
timeout 900 python bench.py --size 128 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench128_ring.json 2> gpurun_out/bench128.err; echo "ring rc=$?"; summ gpurun_out/bench128_ring.json
GENEO_SOLVE_GENERIC=1 timeout 900 python bench.py --size 128 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench128_gen.json 2> gpurun_out/bench128.err; echo "generic rc=$?"; summ gpurun_out/bench128_gen.json
