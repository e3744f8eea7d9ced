#!/bin/bash
# round 2: 8 x B200 -- C3 (heat 400^3 = 64 M DOFs, 64 box subdomains, ASM,1 + CG) and the laplacian weak-scaling point
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
NG=${NG:-8}
for KIND in ${KINDS:-heat laplacian}; do
timeout ${BT:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29871 bench.py --gpus $NG --steps 1 --warmup 1 --kind $KIND ${BENCH_ARGS} > gpurun_out/r2_bench_${KIND}_n${NG}.json 2> gpurun_out/r2_bench_${KIND}_n${NG}.err; echo "bench $KIND rc $?"
grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/r2_bench_${KIND}_n${NG}.err | tail -8
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_${KIND}_n${NG}.json").read().strip().splitlines()[-1])
    print("$KIND n_gpus", d["n_gpus"], d["config"]["workload"]); print(" ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", {k: round(d["e2e"][k],2) for k in ("seconds","symbolic_s","ordering_reuse_s","numeric_s","upload_s","gen_s","part_decomp_s")})
    print(" roofline", d["roofline"]["frac"], "factor TF", d["roofline_factorization"]["achieved"], "its", d["detail"]["iterations"], "dimE", d["detail"]["dimE"], d["detail"]["nev_min_max"], "err", d["detail"]["max_rel_err_vs_1..N"], "iter_s", d["detail"]["iter_s"])
except Exception as e:
    print("no bench line", e)
PY
done
nvidia-smi --query-gpu=index,memory.used --format=csv | head -3
