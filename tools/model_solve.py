"""Host-only model of the level-scheduled solve: per level-phase bytes / item counts from the real symbolic analysis."""
import sys, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, '/root/repo')
import geneo4petsc_b200 as g
from geneo4petsc_b200.api import Symbolic

s = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nsub = int(sys.argv[2]) if len(sys.argv) > 2 else 8
def lap3(s):
    I = sp.identity(s, format='csr'); T = sp.diags([-1, 2, -1], [-1, 0, 1], shape=(s, s), format='csr')
    return (sp.kron(sp.kron(T, I), I) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(I, I), T)).tocsr()
t = time.time(); S = Symbolic(lap3(s)); print('symbolic %.1fs' % (time.time() - t), S.info)
F = S.fronts  # col0,k,h,parent,level,chain,nchild,rowOff,lOff,uOff,wOff,relOff,ld
k, h, lev = F[:, 1], F[:, 2], F[:, 4]
nl = S.info['nlevels']
bytes_f = 8.0 * h * k
BW = 6.5e12
tot = 0; tstream = 0; rows = []
for l in range(nl):
    m = lev == l
    b = bytes_f[m].sum() * nsub
    items = (np.ceil(h[m] / 64) * np.ceil(k[m] / 32)).sum() * nsub
    rows.append((l, m.sum(), b, items, h[m].max()))
print('levels', nl, 'factor GB/sub', bytes_f.sum() / 1e9)
nw = 148 * 16
for lat in (3e-6, 6e-6, 10e-6):
    T = 0
    for (l, nf, b, items, hm) in rows:
        T += 2 * (lat + b / BW)   # fwd + bwd
    print('lat %.0f us: model %.2f ms  (stream only %.2f ms)' % (lat * 1e6, T * 1e3, 2 * sum(r[2] for r in rows) / BW * 1e3))
cum = 0
print('level nfronts MB items maxh')
for r in rows[::max(1, nl // 60)]:
    print(r[0], r[1], '%.2f' % (r[2] / 1e6), int(r[3]), r[4])
small = sum(1 for r in rows if r[3] < nw)
print('levels with < %d items: %d of %d;  bytes in them %.2f GB of %.2f GB' % (nw, small, nl, sum(r[2] for r in rows if r[3] < nw) / 1e9, sum(r[2] for r in rows) / 1e9))
