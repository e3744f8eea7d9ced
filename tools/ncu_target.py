"""Small, fixed workload for ncu captures: 3-D Laplacian S^3, 8 subdomains, setup, then (inside cudaProfilerStart/Stop)
a few PC applies.  Run under `ncu --profile-from-start off` so that only the applies are seen."""
import sys
sys.path.insert(0, ".")
import torch
import geneo4petsc_b200 as g
S = int(sys.argv[1]) if len(sys.argv) > 1 else 48
p = g.Problem().generate("laplacian", "--dim 3 --size %d --inpEps 0.0001" % S).decompose(8, True, 0)
pc = g.GeneoPC(["-geneo_lvl", "ASM,1"]).setup(p)
n = p.sizes()["nb_node"]
x = torch.randn(n, dtype=torch.float64, device="cuda")
y = torch.zeros_like(x)
pc.apply_device(x.data_ptr(), y.data_ptr())
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(3):
    pc.apply_device(x.data_ptr(), y.data_ptr())
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", n, float(y.norm()), pc.stats()["trisolve_bytes"])
