#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
GENEO_HOSTPROF=1 timeout 800 python tools/profile_refactor.py 160 > gpurun_out/hostprof_160.log 2>&1; tail -22 gpurun_out/hostprof_160.log | cut -c1-150
