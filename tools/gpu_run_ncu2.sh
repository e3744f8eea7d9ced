#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_schur -s 55 -c 4 -o gpurun_out/prof_schur_128 -f python tools/ncu_target.py 128 > gpurun_out/ncu_schur_128.log 2>&1; tail -2 gpurun_out/ncu_schur_128.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_solve_ringILi8 -s 20 -c 1 -o gpurun_out/prof_ring8_128 -f python tools/ncu_target.py 128 > gpurun_out/ncu_ring8_128.log 2>&1; tail -2 gpurun_out/ncu_ring8_128.log
ls -la gpurun_out/*.ncu-rep
