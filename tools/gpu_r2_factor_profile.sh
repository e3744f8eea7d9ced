#!/bin/bash
# round 2: where does the factorization time go?  (1) tile-kernel microbenchmarks, (2) ncu launch list of ONE 100^3
# factorization (gpu__time_duration per launch), (3) ncu --set full of the biggest k_schur / k_schur2 launches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
S=${S:-100}
python - <<PY > gpurun_out/r2_microbench.log 2>&1
import sys; sys.path.insert(0, ".")
import geneo4petsc_b200 as g
for n in (1024, 2048, 4096, 8192):
    print("schur shape n=%d K=128: %.2f TFLOP/s, %.3f ms" % ((n,) + g.microbench(3, n, 5)))
for n in (2048, 4096):
    print("square n=%d: %.2f TFLOP/s" % (n, g.microbench(0, n, 3)[0]))
PY
cat gpurun_out/r2_microbench.log
timeout 300 python tools/factor_target.py $S 2 > gpurun_out/r2_factor_plain.log 2>&1; tail -1 gpurun_out/r2_factor_plain.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_factor_launches_${S}.csv python tools/factor_target.py $S 1 > gpurun_out/r2_factor_ncu.log 2>&1; tail -2 gpurun_out/r2_factor_ncu.log
wc -l gpurun_out/r2_factor_launches_${S}.csv
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_schur -s ${SKIP:-300} -c 6 -o gpurun_out/r2_prof_schur_${S} -f python tools/factor_target.py $S 1 > gpurun_out/r2_factor_ncu_full.log 2>&1; tail -2 gpurun_out/r2_factor_ncu_full.log
ls -la gpurun_out/*.ncu-rep 2>/dev/null
