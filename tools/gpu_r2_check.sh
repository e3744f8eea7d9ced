#!/bin/bash
# round 2: GPU parity suite + default bench line (box partition); optional extra args for bench in $BENCH_ARGS
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
TAG=${TAG:-r2}
timeout ${PYTEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q ${PYTEST_ARGS} > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -25 gpurun_out/${TAG}_pytest_gpu.log
if [ -z "$SKIP_BENCH" ]; then
timeout 900 python bench.py --steps 2 --warmup 3 ${BENCH_ARGS} > gpurun_out/${TAG}_bench_box.json 2> gpurun_out/${TAG}_bench_box.err; echo "bench rc $?"
tail -c 2500 gpurun_out/${TAG}_bench_box.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_box.json").read().strip().splitlines()[-1])
    print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["seconds"], "sym", d["e2e"]["symbolic_s"], "ord", d["e2e"]["ordering_reuse_s"], "num", d["e2e"]["numeric_s"])
    print("roofline", d["roofline"]["frac"], "factor TF", d["roofline_factorization"]["achieved"], d["roofline_factorization"]["in_step"]["pipeline_span_s_per_step"])
    print("phases", d["detail"]["numeric_phases_s_rank0"], "its", d["detail"]["iterations"], "dimE", d["detail"]["dimE"])
    print("parity", d.get("parity"))
except Exception as e:
    print("no bench line", e)
PY
fi
