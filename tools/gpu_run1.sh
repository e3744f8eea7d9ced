#!/bin/bash
# round-1 GPU pass: parity tests, smoke, bench at two sizes, ncu launch list + full capture of the solve kernel
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; nvidia-smi --query-gpu=name,memory.total --format=csv >> gpurun_out/host.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --size 128 --steps 1 --warmup 1 > gpurun_out/bench128.json 2> gpurun_out/bench128.err; echo "bench128 rc=$?"; cat gpurun_out/bench128.json; tail -5 gpurun_out/bench128.err
timeout 900 python bench.py --size 200 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench200.json 2> gpurun_out/bench200.err; echo "bench200 rc=$?"; cat gpurun_out/bench200.json; tail -5 gpurun_out/bench200.err
CMD="python bench.py --size 96 --steps 1 --warmup 0 --e2e-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/plain96.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/plain96b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_solve_forest -s 30 -c 2 -o gpurun_out/prof_forest_r1 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
