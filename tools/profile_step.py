"""Per-kernel table of ONE step (numeric re-setup + Krylov solve) on the bench workload (box partition), taken with the
in-library launch profiler: GENEO_PROFILE=1 python tools/profile_step.py KIND SIZE OUT.csv   (sequential numeric path:
the profiler's events live on one stream)."""
import os, sys, time
sys.path.insert(0, '.')
import torch
import geneo4petsc_b200 as g
from geneo4petsc_b200 import dist
kind = sys.argv[1] if len(sys.argv) > 1 else "laplacian"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 200
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/profile_step_%s_%d.csv" % (kind, size)
args = "--dim 3 --size %d --inpEps 0.0001" % size + (" --kappa 100. minmax --lbd 1. --dt 0.1" if kind == "heat" else "")
prob = g.Problem()
K, rg, sub_rank = dist.box_grid(1, 8)
dist.generate_boxed(prob, kind, args, K)
dist.decompose_owned(prob, 8, sub_rank, 0, True, 0)
pc = g.GeneoPC(["-geneo_lvl", "ASM,1", "-geneo_tau", "0.1"] + os.environ.get("GENEO_EXTRA_OPTS", "").split()).setup(prob)
n = size ** 3
b = torch.zeros(n, dtype=torch.float64, device="cuda"); x = torch.zeros_like(b)
ones = torch.arange(1, n + 1, dtype=torch.float64, device="cuda")
pc.mult_device(ones.data_ptr(), b.data_ptr())
pc.refactor()
torch.cuda.synchronize()
g.profile_dump("/tmp/discard.csv")
t = time.time(); pc.refactor(); torch.cuda.synchronize(); t1 = time.time()
r = pc.ksp_solve_device(b.data_ptr(), x.data_ptr(), ksp="cg", rtol=1e-5, atol=1e-50, restart=30)
torch.cuda.synchronize(); t2 = time.time()
g.profile_dump(out)
print("eig steps/dim/nev per subdomain:", [(pc.sub_info(i)["eigSteps"], pc.sub_info(i)["eigDim"], pc.sub_info(i)["nev"]) for i in range(8)])
print("refactor %.3f s, solve %.3f s (%d its)" % (t1 - t, t2 - t1, r["its"]), {k: round(v, 3) for k, v in pc.timers().items() if k.startswith("lvl") and v > 0})
