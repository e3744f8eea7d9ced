#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --size 128 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench128.json 2> gpurun_out/bench128.err; echo "bench128 rc=$?"; cat gpurun_out/bench128.json; tail -5 gpurun_out/bench128.err
GENEO_PROFILE=1 GENEO_PROFILE_OUT=gpurun_out/profile_sites_200.csv timeout 900 python bench.py --size 200 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench200.json 2> gpurun_out/bench200.err; echo "bench200 rc=$?"; cat gpurun_out/bench200.json; tail -5 gpurun_out/bench200.err
python tools/profile_report.py gpurun_out/profile_sites_200.csv > gpurun_out/profile_200.txt; head -16 gpurun_out/profile_200.txt
