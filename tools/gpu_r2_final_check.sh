#!/bin/bash
# final check of round 2: the whole GPU suite, then a short bench line (3 warm-up + 2 timed steps, one fresh e2e setup)
mkdir -p gpurun_out
timeout 210 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_pytest.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/r2_final_pytest.log
timeout 110 python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-budget 8 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
echo "bench rc $?"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r2_final_bench.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("ms_per_step", d["ms_per_step"], "e2e", e["seconds"], {k: round(v, 2) for k, v in e.items() if k.endswith("_s")}, "its", d["detail"]["iterations"], "parity", d.get("parity", {}).get("its_gpu"), d.get("parity", {}).get("its_cpu"))
except Exception as ex:
    print("no bench line:", ex)
P
