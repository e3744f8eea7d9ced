#!/bin/bash
# round 2: new GPU tests, heat N=1 after the Rayleigh-Ritz shortcut, graph C4 at N=1, ncu of the PC-apply kernel on the bench workload
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity2.py tests/test_adapter.py -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc $?"; tail -8 gpurun_out/r2e_pytest.log
bash tools/gpu_r2_heat_prof.sh
timeout 1500 python bench.py --kind graph --subs-per-gpu 64 --steps 1 --warmup 1 --no-cpu-baseline ${GRAPH_ARGS} > gpurun_out/r2_graph_n1.json 2> gpurun_out/r2_graph_n1.err; echo "graph rc $?"; tail -c 600 gpurun_out/r2_graph_n1.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_graph_n1.json").read().strip().splitlines()[-1])
    print("graph", d["config"]["workload"]); print(" ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", {k: d["e2e"][k] for k in ("seconds","symbolic_s","numeric_s","gen_s","part_decomp_s")}, "its", d["detail"]["iterations"], "dimE", d["detail"]["dimE"], d["detail"]["nev_min_max"], "err", d["detail"]["max_rel_err_vs_1..N"], "pc_apply", d["detail"]["pc_apply_rank0"], "spmv", d["detail"]["spmv_rank0"])
except Exception as e:
    print("no graph line", e)
PY
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base mangled -k regex:k_solve_ringILi1 -c 1 -o gpurun_out/r2_prof_ring1_box200 -f python tools/ncu_target_box.py 200 > gpurun_out/r2_ncu_ring1.log 2>&1; tail -2 gpurun_out/r2_ncu_ring1.log
