#!/bin/bash
# BASELINE configs[4]: GenEO threshold sweep on the weak-scaling unit (200^3 per GPU): coarse-space size vs iterations vs time
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for TAU in ${TAUS:-0.05 0.2 0.4}; do
timeout 600 python bench.py --tau $TAU --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_tau_${TAU}.json 2> gpurun_out/r2_tau_${TAU}.err; echo "tau $TAU rc $?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_tau_${TAU}.json").read().strip().splitlines()[-1])
    print("tau ${TAU}: ms_per_step %.0f  its %d  dimE %d  nev %s  e2e %.2f s  iter_s %.3f  phases %s" % (d["ms_per_step"], d["detail"]["iterations"], d["detail"]["dimE"], d["detail"]["nev_min_max"], d["e2e"]["seconds"], d["detail"]["iter_s"], {k: round(v,2) for k,v in d["detail"]["numeric_phases_s_rank0"].items()}))
except Exception as e:
    print("no line", e)
PY
done
