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


def parse_driver_log(text):
    """Parse a log of the geneo4PETSc driver (reference or this repo's CLI): local matrices, B, X, INFO lines --
    the same fields tests/golden/make_dummy_golden.py extracts from the reference's .ref files."""
    g = {"mats": [], "b": [], "x": [], "info": []}
    mode, cur = None, None
    for ln in text.splitlines():
        if ln.startswith("The matrix A is"):
            mode = "A"; continue
        if ln.startswith("The vector B is"):
            mode = "B"; continue
        if ln.startswith("The solution X is"):
            mode = "X"; continue
        if ln.startswith("INFO:"):
            mode = None; g["info"].append(ln); continue
        if mode == "A":
            if "type: seqaij" in ln or "type: mpiaij" in ln:
                cur = []; g["mats"].append(cur)
            m = re.match(r"row (\d+):(.*)", ln)
            if m and cur is not None:
                ent = [[int(a), float(b)] for a, b in re.findall(r"\((\d+), ([-0-9.e+]+)\)", m.group(2))]
                cur.append([int(m.group(1)), ent])
        elif mode in ("B", "X"):
            try:
                g["b" if mode == "B" else "x"].append(float(ln))
            except ValueError:
                pass
    return g


def driver_command(name, g, inputs_dir):
    """The command line of tst/dummy/dummy.sh for golden `name` (mpirun -n 2 becomes --nbPart 2)."""
    cfg = golden_config(g)
    cmd = ["--inpFileA", str(inputs_dir / (g["input"] + ".inp"))]
    if g["input"] == "identity":
        cmd += ["--inpFileB", str(inputs_dir / "B.inp")]
    if g["input"] == "tridiag":
        cmd += ["--inpEps", "1.", "-geneo_cut", "10"]
    cmd += ["-pc_type", "geneo", "-geneo_lvl", "%s,%s" % (cfg["lvl1"], cfg["lvl2"])]
    if cfg["overlap"]:
        cmd += ["--addOverlap", "1"]
    if cfg["offload"]:
        cmd += ["-geneo_offload"]
    cmd += ["--debug", "log", "--verbose", "2", "-geneo_chk", "log", "-geneo_dbg", "log,2", "--shortRes", "-ksp_atol", "1.e-12",
            "-ksp_rtol", "1.e-12", "-options_left", "no", "--metisDual" if cfg["dual"] else "--metisNodal", "--nbPart", "2"]
    return cmd
