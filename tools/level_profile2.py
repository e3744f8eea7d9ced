"""A/B of ring-kernel tuning variants on ONE decomposition: python tools/level_profile2.py SIZE"""
import os, sys, time
import numpy as np
sys.path.insert(0, '.')
import geneo4petsc_b200 as g
size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
t = time.time()
prob = g.Problem().generate("laplacian", "--dim 3 --size %d --inpEps 0.0001" % size).decompose(8, True, 0)
print("decomposition %.1fs" % (time.time() - t), flush=True)
for var in ("0", "1"):
    os.environ["GENEO_RING_VAR"] = var
    t = time.time()
    pc = g.GeneoPC(["-geneo_lvl", "ASM,0"]).setup(prob)
    for rep in range(3):
        us, by, it = pc.level_profile()
    nl = len(us) // 2
    print("VAR=%s setup %.1fs: total %.3f ms, %.1f GB/s (fwd %.0f, bwd %.0f GB/s); phases<8us: %d (%.3f ms)" % (
        var, time.time() - t, us.sum() / 1e3, by.sum() / us.sum() / 1e3, by[:nl].sum() / us[:nl].sum() / 1e3,
        by[nl:].sum() / us[nl:].sum() / 1e3, (us < 8).sum(), us[us < 8].sum() / 1e3), flush=True)
    if var == "0":
        with open("gpurun_out/levels_%d.txt" % size, "w") as f:
            f.write("phase kind lvl us MB items GB/s\n")
            for p in range(len(us)):
                kind, l = ("F", p) if p < nl else ("B", 2 * nl - 1 - p)
                f.write("%d %s %d %.1f %.2f %d %.0f\n" % (p, kind, l, us[p], by[p] / 1e6, it[p], by[p] / max(us[p], 1e-9) / 1e3))
    del pc
