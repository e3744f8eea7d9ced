"""numpy emulation of the device numeric phase of the block LDL^T solver, driven by the EXACT host symbolic structures
the CUDA kernels consume (geneo4petsc_b200/csrc/symbolic.cpp).  Test infrastructure: it validates scatter maps,
relative indices, levels and arena offsets on a CPU-only box."""
import numpy as np

(F_COL0, F_K, F_H, F_PARENT, F_LEVEL, F_CHAIN, F_NCHILD, F_ROWOFF, F_LOFF, F_UOFF, F_WOFF, F_RELOFF, F_LD, F_ULD, F_UARENA,
 F_INPLACE, F_PAIR) = range(17)


def _umat(arenas, fr, f):
    """View of the update matrix of front f: m x m inside arena uArena with leading dimension uLd (symbolic.hpp:Front)."""
    m, ld, off = fr[f, F_H] - fr[f, F_K], fr[f, F_ULD], fr[f, F_UOFF]
    a = arenas[fr[f, F_UARENA]]
    assert off >= 0 and off + (m - 1) * ld + m <= len(a), "update matrix outside its arena"
    return np.lib.stride_tricks.as_strided(a[off:], shape=(m, m), strides=(a.itemsize, ld * a.itemsize))


def _panel(L, fr, f):
    """View of panel f: h x k, column-major with leading dimension ld (h rounded up to even)."""
    k, h, ld = fr[f, F_K], fr[f, F_H], fr[f, F_LD]
    return L[fr[f, F_LOFF]: fr[f, F_LOFF] + ld * k].reshape(ld, k, order="F")[:h, :]


def factorize(sym, vals):
    """vals: CSR values of the input matrix (same pattern/order as given to Symbolic).  Returns (L, neg)."""
    fr = sym.fronts
    L = np.zeros(sym.info["lSize"])
    L[sym.asm_dst] = vals[sym.asm_src]
    # ping-pong arenas 0/1 and the chain arena 2, poisoned with NaN: whatever is read must have been zeroed or written
    arenas = [np.full(max(1, sym.info["uArena"]), np.nan), np.full(max(1, sym.info["uArena"]), np.nan),
              np.full(max(1, sym.info["cArena"]), np.nan)]
    neg = 0
    Wbuf = [np.full(max(1, sym.info["wArena"]), np.nan), np.full(max(1, sym.info["wArena"]), np.nan)]
    order = np.argsort(fr[:, F_LEVEL], kind="stable")
    levels = fr[:, F_LEVEL]
    for lvl in range(sym.info["nlevels"]):
        arenas[lvl & 1][:] = 0.0
        for f in order[levels[order] == lvl]:  # chain blocks are zeroed where they are born
            if fr[f, F_H] > fr[f, F_K] and fr[f, F_UARENA] == 2 and not fr[f, F_INPLACE]:
                m = fr[f, F_H] - fr[f, F_K]
                arenas[2][fr[f, F_UOFF]: fr[f, F_UOFF] + m * m] = 0.0
        W = Wbuf[lvl & 1]  # the unscaled panels of a level live in the scratch of its parity: the second panel of a pair reads
        Wprev = Wbuf[(lvl & 1) ^ 1]  # the first one's a level later
        W[:] = np.nan
        if lvl > 0:  # extend-add children (level lvl-1) into their parents
            for c in order[levels[order] == lvl - 1]:
                k, h, par = fr[c, F_K], fr[c, F_H], fr[c, F_PARENT]
                m = h - k
                if m == 0:
                    continue
                assert par >= 0 and fr[par, F_LEVEL] == lvl
                pk, ph = fr[par, F_K], fr[par, F_H]
                pm = ph - pk
                U = _umat(arenas, fr, c)
                rel = np.arange(m) if fr[c, F_RELOFF] < 0 else sym.rel[fr[c, F_RELOFF]: fr[c, F_RELOFF] + m]
                Pp = _panel(L, fr, par)
                Up = _umat(arenas, fr, par) if pm > 0 else None
                inplace = bool(fr[par, F_INPLACE])
                if inplace:  # the parent's update matrix IS the trailing block of the child's
                    assert fr[c, F_CHAIN] == 1 and fr[par, F_NCHILD] == 1 and fr[c, F_RELOFF] < 0
                    assert fr[par, F_UARENA] == fr[c, F_UARENA] and fr[par, F_ULD] == fr[c, F_ULD]
                    assert fr[par, F_UOFF] == fr[c, F_UOFF] + pk * (fr[c, F_ULD] + 1) and pm == m - pk
                for cc in range(m):
                    pc = rel[cc]
                    rr = np.arange(cc, m)
                    pr = rel[rr]
                    if pc < pk:
                        Pp[pr, pc] += U[rr, cc]
                    elif not inplace:
                        Up[pr - pk, pc - pk] += U[rr, cc]
        for f in order[levels[order] == lvl]:
            k, h = fr[f, F_K], fr[f, F_H]
            m = h - k
            P = _panel(L, fr, f)
            F11 = np.tril(P[:k, :k]) + np.tril(P[:k, :k], -1).T
            # symmetric sweep (pivot signs = inertia)
            a = F11.copy()
            for p in range(k):
                d = a[p, p]
                if d < 0:
                    neg += 1
                col = a[:, p].copy()
                a -= np.outer(col, col) / d
                a[:, p] = col / d
                a[p, :] = col / d
                a[p, p] = -1.0 / d
            Dinv = -a
            P[:k, :k] = Dinv
            if m > 0:
                F21 = P[k:, :].copy()
                Wf = W[fr[f, F_WOFF]: fr[f, F_WOFF] + m * k].reshape(m, k, order="F")
                Wf[:, :] = F21
                P[k:, :] = F21 @ Dinv
                U = _umat(arenas, fr, f)
                il = np.tril_indices(m)
                if fr[f, F_PAIR] == 1:  # first of a pair: only the strip the next panel assembles (its k' pivot columns)
                    assert fr[f + 1, F_PAIR] == 2 and fr[f + 1, F_INPLACE] == 1 and fr[f, F_PARENT] == f + 1
                    kn = fr[f + 1, F_K]
                    upd = P[k:, :] @ F21[:kn, :].T
                    strip = il[1] < kn
                    U[il[0][strip], il[1][strip]] -= upd[il[0][strip], il[1][strip]]
                elif fr[f, F_PAIR] == 2:  # second of a pair: both panels at once on the trailing block (K = k_prev + k)
                    pf = f - 1
                    assert fr[pf, F_PAIR] == 1 and fr[pf, F_LEVEL] == lvl - 1
                    pk, pm = fr[pf, F_K], fr[pf, F_H] - fr[pf, F_K]
                    Pp = _panel(L, fr, pf)
                    Wp = Wprev[fr[pf, F_WOFF]: fr[pf, F_WOFF] + pm * pk].reshape(pm, pk, order="F")
                    upd = P[k:, :] @ F21.T + Pp[pk + k:, :] @ Wp[k:, :].T
                    U[il] -= upd[il]
                else:
                    upd = P[k:, :] @ F21.T
                    U[il] -= upd[il]  # the device writes whole 64x64 tiles; only the lower triangle is ever read
    return L, neg


def solve(sym, L, b):
    """Solve in the ORIGINAL ordering using the permuted block LDL^T factor."""
    fr = sym.fronts
    x = b[sym.perm].astype(float).copy()
    order = np.argsort(fr[:, F_LEVEL], kind="stable")
    for f in order:  # forward (levels ascending)
        k, h = fr[f, F_K], fr[f, F_H]
        if h == k:
            continue
        P = _panel(L, fr, f)
        rows = sym.row_idx[fr[f, F_ROWOFF]: fr[f, F_ROWOFF] + h]
        np.subtract.at(x, rows[k:], P[k:, :] @ x[rows[:k]])
    y = np.zeros_like(x)
    for f in range(len(fr)):
        k, h = fr[f, F_K], fr[f, F_H]
        P = _panel(L, fr, f)
        rows = sym.row_idx[fr[f, F_ROWOFF]: fr[f, F_ROWOFF] + k]
        assert np.array_equal(rows, np.arange(fr[f, F_COL0], fr[f, F_COL0] + k))
        y[rows] = P[:k, :] @ x[rows]
    for f in order[::-1]:  # backward (levels descending)
        k, h = fr[f, F_K], fr[f, F_H]
        if h == k:
            continue
        P = _panel(L, fr, f)
        rows = sym.row_idx[fr[f, F_ROWOFF]: fr[f, F_ROWOFF] + h]
        y[rows[:k]] -= P[k:, :].T @ y[rows[k:]]
    out = np.zeros_like(y)
    out[sym.perm] = y
    return out
