"""Shared helpers for golden / parity tests (test infrastructure)."""
import re

PC_RE = re.compile(r"geneo(ASM|SORAS)(0|1|H1|E1|2|H2|E2)$")


def golden_config(g):
    """Translate a tst/dummy golden's name fields into (lvl1, lvl2, dual, overlap, offload), or None for bjacobi."""
    m = PC_RE.match(g["pc"])
    if not m:
        return None
    return dict(lvl1=m.group(1), lvl2=m.group(2), dual=(g["metis"] == "dual"),
                overlap=1 if "overlap1" in g["opt"] else 0, offload="offload" in g["opt"])


def dense_from_golden(rows, n):
    import numpy as np
    a = np.zeros((n, n))
    for r, ent in rows:
        for c, v in ent:
            a[r, c] = v
    return a
