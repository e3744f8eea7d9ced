#!/bin/bash
# round 2: multi-GPU parity (2 ranks) + N-GPU bench line; usage: NG=2 bash tools/gpu_r2_dist.sh
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
NG=${NG:-2}
TAG=${TAG:-r2}
if [ -z "$SKIP_TESTS" ]; then
timeout 1200 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/${TAG}_pytest_dist.log 2>&1; echo "pytest rc $?" >> gpurun_out/${TAG}_pytest_dist.log
tail -15 gpurun_out/${TAG}_pytest_dist.log
fi
for KIND in ${KINDS:-laplacian}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29871 bench.py --gpus $NG --steps ${STEPS:-2} --warmup ${WARMUP:-2} --kind $KIND ${BENCH_ARGS} > gpurun_out/${TAG}_bench_${KIND}_n${NG}.json 2> gpurun_out/${TAG}_bench_${KIND}_n${NG}.err; echo "bench $KIND rc $?"
tail -c 1500 gpurun_out/${TAG}_bench_${KIND}_n${NG}.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${KIND}_n${NG}.json").read().strip().splitlines()[-1])
    print("$KIND n_gpus", d["n_gpus"], "ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", {k: d["e2e"][k] for k in ("seconds","symbolic_s","ordering_reuse_s","numeric_s","upload_s","gen_s","part_decomp_s")})
    print("roofline", d["roofline"]["frac"], "factor TF", d["roofline_factorization"]["achieved"], "its", d["detail"]["iterations"], "dimE", d["detail"]["dimE"], d["detail"]["nev_min_max"], "err", d["detail"]["max_rel_err_vs_1..N"])
except Exception as e:
    print("no bench line", e)
PY
done
