"""CPU: stage stopwatches + digest of the per-subdomain host preparation (geneo_host_prepare_probe) on an n^3 Q1 box.

python tools/host_prepare_probe.py [edge] [--metis] [--seven]   (inherited reference-box ordering, 27-point pattern by default)
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from geneo4petsc_b200 import api  # noqa: E402


def q1_box(n, seed=0, seven=False):
    b = sp.diags([np.ones(n - 1), np.ones(n), np.ones(n - 1)], [-1, 0, 1], format="csr")
    i = sp.identity(n, format="csr")
    if seven:  # 7-point pattern (the bench's laplacian / heat generators: 2-node elements along the axes)
        pat = (sp.kron(sp.kron(b, i), i) + sp.kron(sp.kron(i, b), i) + sp.kron(sp.kron(i, i), b)).tocsr()
    else:      # 27-point pattern (Q1 hexahedra)
        pat = sp.kron(sp.kron(b, b), b).tocsr()
    rng = np.random.default_rng(seed)
    w = sp.triu(pat, 1).tocoo()
    v = -rng.uniform(0.5, 1.5, w.nnz)
    off = sp.coo_matrix((v, (w.row, w.col)), shape=pat.shape)
    a = (off + off.T).tocsr()
    a = a + sp.diags(-np.asarray(a.sum(axis=1)).ravel() + 1.0)
    a = a.tocsr()
    a.sort_indices()
    return a


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 64
    seven = "--seven" in sys.argv
    a = q1_box(n, seven=seven)
    perm = None
    if "--metis" not in sys.argv:
        st = [(i, j, k) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1) if (i, j, k) > (0, 0, 0)]
        if seven:
            st = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
        t = time.time()
        rank = api.box_ordering((n, n, n), stencil=st, threads=os.cpu_count())
        print("reference ordering %.2f s" % (time.time() - t))
        perm = np.argsort(rank).astype(np.int32)
    for helper in (False, True):
        t = time.time()
        sec, dig = api.host_prepare_probe(a, perm, helper=helper)
        print("helper %d wall %.2f s;" % (helper, time.time() - t), end=" ")
        print("edge %d nnz %d: analysis %.3f s, permuted values %.3f s, work lists %.3f s, digest %016x" %
              (n, a.nnz, sec[0], sec[1], sec[2], dig))
