"""Workload for ncu launch lists of the factorization: ONE box subdomain of S^3 nodes (GenEO level 1 only: a single block
LDL^T factorization per setup), sequential path.  usage: python tools/factor_target.py S [refactors]"""
import os
import sys
sys.path.insert(0, ".")
os.environ["GENEO_PIPELINE"] = "0"
import torch
import geneo4petsc_b200 as g
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p = g.Problem().generate("laplacian", "--dim 3 --size %d --inpEps 0.0001" % S).decompose(1, True, 0)
pc = g.GeneoPC(["-geneo_lvl", "ASM,0"]).setup(p)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(R):
    pc.refactor()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
st = pc.stats()
fs = pc.factor_stats()
print("ok n %d factor_flops %.4e factor_bytes %.4e seconds %.4f TFLOPs %.2f" % (S ** 3, st["factor_flops"], st["factor_bytes"], fs["seconds"],
                                                                                fs["flops"] / fs["seconds"] / 1e12))
