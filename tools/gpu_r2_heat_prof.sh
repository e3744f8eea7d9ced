#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
GENEO_HOSTPROF=1 timeout 900 python bench.py --kind heat --steps 1 --warmup 1 --no-cpu-baseline --size ${HS:-200} > gpurun_out/r2_heat_n1.json 2> gpurun_out/r2_heat_n1.err; echo "heat rc $?"
grep HOSTPROF gpurun_out/r2_heat_n1.err | tail -45
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_heat_n1.json").read().strip().splitlines()[-1])
print("heat ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["seconds"], "its", d["detail"]["iterations"], "dimE", d["detail"]["dimE"], d["detail"]["nev_min_max"], d["detail"]["numeric_phases_s_rank0"], "iter_s", d["detail"]["iter_s"])
PY
