"""Scaling exploration (not a test): one setup + solve at a given grid size, prints where the time goes."""
import argparse, sys, time, json
sys.path.insert(0, '.')
import numpy as np
import geneo4petsc_b200 as g

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=64)
ap.add_argument("--dim", type=int, default=3)
ap.add_argument("--nparts", type=int, default=8)
ap.add_argument("--ksp", default="cg")
ap.add_argument("--lvl", default="ASM,1")
ap.add_argument("--tau", default="0.1")
ap.add_argument("--rtol", type=float, default=1e-5)
ap.add_argument("--nodal", action="store_true")
ap.add_argument("--overlap", type=int, default=0)
ap.add_argument("--box", action="store_true", help="explicit box partition instead of METIS")
ap.add_argument("--extra", default="")
a = ap.parse_args()
t0 = time.time()
p = g.Problem().generate("laplacian", "--dim %d --size %d --inpEps 0.0001" % (a.dim, a.size))
t1 = time.time()
if a.box:
    s = p.sizes()
    ep, ei, em = p.mesh()
    n = a.size
    first = ei[ep[:-1]]
    k = round(a.nparts ** (1 / 3))
    i, j, l = first % n, (first // n) % n, first // (n * n)
    part = ((i * k) // n) + k * ((j * k) // n) + k * k * ((l * k) // n)
    p.decompose(a.nparts, True, a.overlap, elem_part=part.astype(np.int32))
else:
    p.decompose(a.nparts, not a.nodal, a.overlap)
t2 = time.time()
print("gen %.2fs part/decomp %.2fs" % (t1 - t0, t2 - t1), p.sizes(), flush=True)
pc = g.GeneoPC(["-geneo_lvl", a.lvl, "-geneo_tau", a.tau] + a.extra.split())
pc.setup(p)
t3 = time.time()
print("setup %.2fs" % (t3 - t2), flush=True)
tm = pc.timers()
print({k: round(v, 3) for k, v in tm.items() if v > 0})
st = pc.stats()
print({k: "%.3e" % v for k, v in st.items()})
print(pc.info())
print([(pc.sub_info(s)["n"], pc.sub_info(s)["nev"], pc.sub_info(s)["eigSteps"], pc.sub_info(s)["eigDim"], pc.sub_info(s)["perturbed"]) for s in range(a.nparts)])
b = pc.make_rhs()
t4 = time.time()
r = pc.ksp_solve(b, ksp=a.ksp, rtol=a.rtol, atol=1e-50)
t5 = time.time()
n = len(b)
err = np.abs(r["x"] - np.arange(1, n + 1.0)).max() / n
print("solve %.3fs its %d %s rnorm %.3e relerr %.2e" % (t5 - t4, r["its"], r["reason_name"], r["rnorm"], err))
import torch
x = torch.randn(n, dtype=torch.float64, device="cuda"); y = torch.zeros_like(x)
for name, fn, bytes_ in (("apply", pc.apply_device, st["apply_bytes"]), ("mult", pc.mult_device, st["spmv_bytes"])):
    for _ in range(3): fn(x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn(x.data_ptr(), y.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%s: %.3f ms  %.1f GB/s (algorithmic)" % (name, ms, bytes_ / ms / 1e6))
print("tri-solve bytes/apply %.3e  factor flops %.3e  L1 factor time -> %.2f TFLOP/s" % (st["trisolve_bytes"], st["factor_flops"], st["factor_flops"] / max(tm["lvl1SetupMinv"], 1e-9) / 1e12))
