#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
GENEO_HOSTPROF=1 timeout 800 python tools/profile_refactor.py 160 > gpurun_out/hostprof_160.log 2>&1; tail -16 gpurun_out/hostprof_160.log
GENEO_PROFILE=1 timeout 900 python tools/profile_refactor.py 160 > gpurun_out/profile_refactor_160.log 2>&1; tail -1 gpurun_out/profile_refactor_160.log | cut -c1-200
python tools/profile_report.py gpurun_out/profile_refactor_160.csv | tee gpurun_out/profile_refactor_160.txt | head -16
