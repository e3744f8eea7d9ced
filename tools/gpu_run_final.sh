#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --gpus 1 --steps 2 --warmup 3 > gpurun_out/bench200_default.json 2> gpurun_out/bench200_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench200_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench200_default.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],d["e2e"]["seconds"],"frac",d["roofline"]["frac"],"launches",d["gpu_launches"], d["clocks"])
print(d["detail"]["numeric_phases_s"], "iter", d["detail"]["iter_s"], "its", d["detail"]["iterations"])
print(d.get("cpu_baseline"))
PY
GENEO_PROFILE=1 timeout 900 python tools/profile_refactor.py 200 gpurun_out/profile_step_200.csv > gpurun_out/profile_step_200.log 2>&1; tail -1 gpurun_out/profile_step_200.log | cut -c1-160
python tools/profile_report.py gpurun_out/profile_step_200.csv > gpurun_out/profile_step_200.txt; head -14 gpurun_out/profile_step_200.txt
