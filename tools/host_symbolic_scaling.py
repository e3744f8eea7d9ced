"""Host-only experiment: how does the symbolic analysis (METIS nested dissection) of 100^3 subdomains scale with threads on
this box?  usage: python tools/host_symbolic_scaling.py SIZE"""
import os, sys, time, threading
import numpy as np, scipy.sparse as sp
sys.path.insert(0, '.')
from geneo4petsc_b200.api import Symbolic
s = int(sys.argv[1]) if len(sys.argv) > 1 else 100
I = sp.identity(s, format='csr'); T = sp.diags([-1., 2., -1.], [-1, 0, 1], shape=(s, s), format='csr')
A = (sp.kron(sp.kron(T, I), I) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(I, I), T)).tocsr()
print("cores", os.cpu_count(), flush=True)
def run(nconc, depth):
    os.environ["GENEO_ND_DEPTH"] = str(depth)
    res = []
    def work():
        t = time.time(); S = Symbolic(A); res.append((round(time.time() - t, 1), round(S.info['lSize'] / 1e6)))
    t0 = time.time()
    ths = [threading.Thread(target=work) for _ in range(nconc)]
    [t.start() for t in ths]; [t.join() for t in ths]
    print("concurrent %d depth %d: wall %.1fs -> %.2f s per subdomain  %s" % (nconc, depth, time.time() - t0, (time.time() - t0) / nconc, res[:3]), flush=True)
for nconc, depth in ((1, 0), (1, 2), (1, 3), (1, 4), (2, 3), (4, 2), (8, 1), (8, 0)):
    run(nconc, depth)
