"""Workload for ncu captures of the PC-apply kernel ON THE BENCH WORKLOAD: S^3 grid, 2x2x2 box partition, ASM,1; after the
setup a few PC applies inside cudaProfilerStart/Stop (run under `ncu --profile-from-start off`)."""
import sys
sys.path.insert(0, ".")
import torch
import geneo4petsc_b200 as g
from geneo4petsc_b200 import dist
S = int(sys.argv[1]) if len(sys.argv) > 1 else 200
prob = g.Problem()
K, rg, sub_rank = dist.box_grid(1, 8)
dist.generate_boxed(prob, "laplacian", "--dim 3 --size %d --inpEps 0.0001" % S, K)
dist.decompose_owned(prob, 8, sub_rank, 0, True, 0)
pc = g.GeneoPC(["-geneo_lvl", "ASM,1"]).setup(prob)
n = S ** 3
x = torch.randn(n, dtype=torch.float64, device="cuda")
y = torch.zeros_like(x)
pc.apply_device(x.data_ptr(), y.data_ptr())
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(2):
    pc.apply_device(x.data_ptr(), y.data_ptr())
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", n, float(y.norm()), pc.stats()["trisolve_bytes"])
