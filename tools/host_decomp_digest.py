"""CPU: SHA-256 of everything the decomposition and the rank layouts produce, over a fixed list of small configurations
(regression check when mesh.cpp is reworked: the digests must not move).

python tools/host_decomp_digest.py [out.json]
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import geneo4petsc_b200 as g  # noqa: E402
from geneo4petsc_b200 import dist  # noqa: E402


def digest_problem(prob, nb_part, owned, layout=None):
    h = hashlib.sha256()
    for s in range(nb_part):
        nodes, mult = prob.sub_nodes(s)
        h.update(nodes.tobytes()); h.update(mult.tobytes())
        for q in range(nb_part):
            h.update(prob.sub_intersect(s, q).tobytes())
        if owned is None or owned[s]:
            for which in (0, 1):
                m = prob.sub_matrix(s, which)
                h.update(m.indptr.astype(np.int64).tobytes()); h.update(m.indices.astype(np.int32).tobytes()); h.update(m.data.tobytes())
    if layout is not None:
        h.update(layout.owned.tobytes()); h.update(layout.ghost.tobytes()); h.update(layout.ghost_ptr.tobytes())
        m = layout.matrix()
        h.update(m.indptr.astype(np.int64).tobytes()); h.update(m.indices.astype(np.int32).tobytes()); h.update(m.data.tobytes())
    return h.hexdigest()


def run():
    out = {}
    gen = {"laplacian": "--dim 3 --size %d --inpEps 0.0001", "heat": "--dim 3 --size %d --inpEps 0.0001 --kappa 100. minmax --lbd 1. --dt 0.1",
           "lap2d": "--dim 2 --size %d --inpEps 0.0001"}
    # box partitions, one rank's view
    for world, size, kind in ((1, 14, "laplacian"), (2, 12, "heat"), (8, 8, "laplacian"), (4, 10, "laplacian")):
        edge = int(np.floor((float(size) ** 3 * world) ** (1.0 / 3.0) + 1e-9)) if world > 1 else size
        K, rg, sub_rank = dist.box_grid(world, 8)
        for rank in sorted({0, world // 2, world - 1}):
            lo, hi = dist.keep_region(edge, K, rg, rank, 8) if world > 1 else (None, None)
            prob = g.Problem()
            dist.generate_boxed(prob, kind, gen[kind] % edge, K, lo, hi)
            dist.decompose_owned(prob, len(sub_rank), sub_rank, rank, True, 0)
            lay = dist.Layout(prob, rank, world, sub_rank) if world > 1 else None
            out["box w%d r%d %s %d" % (world, rank, kind, size)] = digest_problem(prob, len(sub_rank), sub_rank == rank, lay)
    # METIS partitions (dual / nodal, overlap 0..2), whole problem on one rank
    for kind, size, nb, dual, ov in (("laplacian", 9, 5, True, 0), ("laplacian", 9, 5, True, 2), ("lap2d", 24, 7, False, 1),
                                      ("heat", 8, 4, False, 0), ("lap2d", 30, 6, True, 1)):
        prob = g.Problem().generate("laplacian" if kind == "lap2d" else kind, gen[kind] % size)
        prob.decompose(nb, dual, ov)
        out["metis %s %d p%d dual%d ov%d" % (kind, size, nb, dual, ov)] = digest_problem(prob, nb, None)
    # METIS parts grouped onto 2 and 3 ranks
    for world, nb, dual, ov in ((2, 6, True, 0), (3, 7, True, 1), (2, 4, False, 1)):
        for rank in range(world):
            prob = g.Problem().generate("laplacian", gen["laplacian"] % 10)
            sub_rank = dist.metis_problem(prob, nb, world, rank, dual, ov)
            lay = dist.Layout(prob, rank, world, sub_rank)
            out["grouped w%d r%d p%d dual%d ov%d" % (world, rank, nb, dual, ov)] = digest_problem(prob, nb, sub_rank == rank, lay)
    # the graph generator
    prob = g.Problem().generate("graph", "--size 2000 --level 3 --weakScaling 1 --noGround --inpEps 0.0001")
    prob.decompose(6, True, 0)
    out["graph 2000 l3 p6"] = digest_problem(prob, 6, None)
    return out


if __name__ == "__main__":
    res = run()
    if len(sys.argv) > 1 and os.path.exists(sys.argv[1]):
        old = json.load(open(sys.argv[1]))
        bad = [k for k in res if old.get(k) != res[k]]
        print("compared %d configurations with %s: %s" % (len(res), sys.argv[1], "IDENTICAL" if not bad else "DIFFERENT: %s" % bad))
        sys.exit(1 if bad else 0)
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)
    for k, v in res.items():
        print(k, v[:16])
