#!/bin/bash
# ncu evidence: (1) k_solve_ring<1> on the bench workload (200^3) -> DRAM traffic; (2) k_schur and k_solve_ring<8> at 128^3
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_solve_ring -c 1 -o gpurun_out/prof_ring1_200 -f python tools/ncu_target.py 200 > gpurun_out/ncu_ring1_200.log 2>&1; tail -2 gpurun_out/ncu_ring1_200.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_schur -s 2600 -c 3 -o gpurun_out/prof_schur_128 -f python tools/ncu_target.py 128 > gpurun_out/ncu_schur_128.log 2>&1; tail -2 gpurun_out/ncu_schur_128.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve_ringILi8 -s 20 -c 1 -o gpurun_out/prof_ring8_128 -f python tools/ncu_target.py 128 > gpurun_out/ncu_ring8_128.log 2>&1; tail -2 gpurun_out/ncu_ring8_128.log
ls -la gpurun_out/*.ncu-rep
