#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
GENEO_PROFILE=1 GENEO_PROFILE_OUT=gpurun_out/profile_sites_200.csv timeout 900 python bench.py --size 200 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench200.json 2> gpurun_out/bench200.err; echo "bench200 rc=$?"; cat gpurun_out/bench200.json; tail -5 gpurun_out/bench200.err
python tools/profile_report.py gpurun_out/profile_sites_200.csv | tee gpurun_out/profile_200.txt
python tools/ncu_target.py 128 > gpurun_out/ncu_plain.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_solve_forest -c 1 -o gpurun_out/prof_forest1_128 python tools/ncu_target.py 128 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -5 gpurun_out/ncu_full.log
