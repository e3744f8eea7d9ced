#!/bin/bash
# First contact with the B200: calibration microbenchmarks + the gpu test-suite, logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt 2>&1
python - > gpurun_out/microbench.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
import geneo4petsc_b200 as g
print("devices", g.device_count())
for n in (1024, 4096, 8192):
    print("dmma gemm", n, g.microbench(0, n, 3))
print("copy GB/s", g.microbench(1, 1 << 28, 10))
import torch, time
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
torch.matmul(a, b); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): torch.matmul(a, b)
e1.record(); torch.cuda.synchronize()
print("cublas dgemm 8192 TFLOP/s", 5 * 2 * 8192**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
PY
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -120 gpurun_out/pytest_gpu.log | cut -c1-200
cat gpurun_out/microbench.log
