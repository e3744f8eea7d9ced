#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -4 gpurun_out/pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --size 200 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_n2_200.json 2> gpurun_out/bench_n2_200.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/bench_n2_200.err; cut -c1-2600 gpurun_out/bench_n2_200.json
