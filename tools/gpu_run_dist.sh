#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -25 gpurun_out/pytest_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --size 64 --steps 1 --warmup 1 > gpurun_out/bench_n2_64.json 2> gpurun_out/bench_n2_64.err; echo "bench n2 rc=$?"; cat gpurun_out/bench_n2_64.json; tail -15 gpurun_out/bench_n2_64.err
