#!/bin/bash
# ncu --set full of ONE k_solve_ring<8> launch on a single 100^3 shift-invert factor (the block solve of the eigen-solver)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
GENEO_PIPELINE=0 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_solve_ringILi8 -s 3 -c 1 -o gpurun_out/r2_prof_ring8_100 -f python tools/lanes_target.py 200 ASM,1 > gpurun_out/r2_ncu_ring8.log 2>&1; tail -3 gpurun_out/r2_ncu_ring8.log
ls -la gpurun_out/r2_prof_ring8_100.ncu-rep
