"""Lane experiment: 8 box subdomains of ~100^3, level 1 only (8 independent LDL^T factorizations per re-setup): how does the
re-setup time change with the number of pipeline lanes?  usage: GENEO_LANES=k python tools/lanes_target.py [size] [lvl]"""
import os
import sys
import time
sys.path.insert(0, ".")
import torch
import geneo4petsc_b200 as g
from geneo4petsc_b200 import dist
S = int(sys.argv[1]) if len(sys.argv) > 1 else 200
lvl = sys.argv[2] if len(sys.argv) > 2 else "ASM,0"
prob = g.Problem()
K, rg, sub_rank = dist.box_grid(1, 8)
dist.generate_boxed(prob, "laplacian", "--dim 3 --size %d --inpEps 0.0001" % S, K)
dist.decompose_owned(prob, 8, sub_rank, 0, True, 0)
pc = g.GeneoPC(["-geneo_lvl", lvl]).setup(prob)
for _ in range(2):
    pc.refactor()
torch.cuda.synchronize()
t = time.perf_counter()
R = 3
for _ in range(R):
    pc.refactor()
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / R
fs = pc.factor_stats()
print("lanes %s pipeline %s lvl %s: refactor %.3f s, factor flops %.3e -> %.2f TFLOP/s, phases %s" % (
    os.environ.get("GENEO_LANES", "default"), os.environ.get("GENEO_PIPELINE", "1"), lvl, dt, fs["flops"], fs["flops"] / dt / 1e12,
    {k: round(v, 3) for k, v in pc.timers().items() if k.startswith("lvl") and "Setup" in k and v > 0}))
