"""Cold and warm setup time of a small problem (48^3, 8 METIS subdomains): the fixed costs of the numeric pipeline."""
import sys, time
sys.path.insert(0, ".")
import torch
import geneo4petsc_b200 as g
S = int(sys.argv[1]) if len(sys.argv) > 1 else 48
p = g.Problem().generate("laplacian", "--dim 3 --size %d --inpEps 0.0001" % S).decompose(8, True, 0)
torch.cuda.synchronize()
for rep in range(2):
    t = time.perf_counter(); pc = g.GeneoPC(["-geneo_lvl", "ASM,1"]).setup(p); torch.cuda.synchronize(); t1 = time.perf_counter()
    pc.refactor(); torch.cuda.synchronize(); t2 = time.perf_counter()
    pc.refactor(); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("rep %d: cold setup %.3f s, refactor %.3f s, refactor %.3f s  %s" % (rep, t1 - t, t2 - t1, t3 - t2, {k: round(v, 3) for k, v in pc.timers().items() if v > 0.0005 and ("Setup" in k or k in ("symbolic", "upload", "numeric", "operator", "setup"))}))
    del pc
