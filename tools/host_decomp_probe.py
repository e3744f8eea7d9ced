"""CPU: wall time of generate + decompose (+ layout) of ONE rank of the box-partitioned bench workload.

python tools/host_decomp_probe.py [world] [rank] [size-per-gpu]   (default 8 0 200: the N = 8 weak-scaling point)
"""
import os
import sys
import time
import types

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import geneo4petsc_b200 as g  # noqa: E402
from geneo4petsc_b200 import dist  # noqa: E402
import bench  # noqa: E402

if __name__ == "__main__":
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    size = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    edge = int(round((size ** 3 * world) ** (1.0 / 3.0) + 1e-9))
    while edge ** 3 > size ** 3 * world:
        edge -= 1
    a = types.SimpleNamespace(subs_per_gpu=8, kind="laplacian", eps=1e-4, graph_level=5)
    t0 = time.time()
    K, rg, sub_rank = dist.box_grid(world, 8)
    lo, hi = dist.keep_region(edge, K, rg, rank, 8) if world > 1 else (None, None)
    prob = g.Problem()
    dist.generate_boxed(prob, "laplacian", bench.gen_args(a, edge), K, lo, hi)
    t1 = time.time()
    dist.decompose_owned(prob, len(sub_rank), sub_rank, rank, True, 0)
    t2 = time.time()
    lay = dist.Layout(prob, rank, world, sub_rank) if world > 1 else None
    t3 = time.time()
    print("world %d rank %d edge %d: generate %.2f s, decompose %.2f s, layout %.2f s" % (world, rank, edge, t1 - t0, t2 - t1, t3 - t2))
