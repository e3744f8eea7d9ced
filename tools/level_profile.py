"""Per-level timing of the PC-apply solve kernel (device timestamps after every grid barrier) on a real decomposition.
usage: python tools/level_profile.py SIZE [NSUB]   -> table on stdout"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import geneo4petsc_b200 as g

size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nsub = int(sys.argv[2]) if len(sys.argv) > 2 else 8
t = time.time()
prob = g.Problem().generate("laplacian", "--dim 3 --size %d --inpEps 0.0001" % size).decompose(nsub, True, 0)
pc = g.GeneoPC(["-geneo_lvl", "ASM,0"]).setup(prob)
print("setup %.1fs" % (time.time() - t))
for rep in range(2):
    us, by, it = pc.level_profile()
nl = len(us) // 2
tot = us.sum()
print("phases %d total %.3f ms  bytes %.3f GB  -> %.1f GB/s" % (len(us), tot / 1e3, by.sum() / 1e9, by.sum() / tot / 1e3))
print("phase  kind lvl      us       MB   items    GB/s   ideal_us")
for p in range(len(us)):
    kind, l = ("F", p) if p < nl else ("B", 2 * nl - 1 - p)
    print("%5d  %s %4d %8.1f %8.2f %7d %7.0f %8.1f" % (p, kind, l, us[p], by[p] / 1e6, it[p], by[p] / max(us[p], 1e-9) / 1e3, by[p] / 6.5e6))
fw, bw = us[:nl].sum(), us[nl:].sum()
print("forward %.3f ms (%.0f GB/s)  backward %.3f ms (%.0f GB/s)" % (fw / 1e3, by[:nl].sum() / fw / 1e3, bw / 1e3, by[nl:].sum() / bw / 1e3))
small = us < 8.0
print("phases under 8 us: %d, %.3f ms total" % (small.sum(), us[small].sum() / 1e3))
