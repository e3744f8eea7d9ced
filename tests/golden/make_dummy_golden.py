#!/usr/bin/env python
"""Parse the reference's 84 tst/dummy/*.ref golden logs (+ the 3 .inp inputs) into one JSON fixture.

Run in the build container (needs /root/reference):  python tests/golden/make_dummy_golden.py
The fixture (tests/golden/dummy_goldens.json) is what travels; tests never read /root/reference.
Each golden pins: per-rank local (Neumann) matrices of the MATIS operator, RHS, solution, the INFO lines
(counts, ksp type / tolerances, PC name string, solver names)  -- see SURVEY.md section 4.
"""
import glob
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/tst/dummy"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dummy_goldens.json")


def parse(path):
    lines = open(path).read().splitlines()
    g = {"mats": [], "b": [], "x": [], "info": []}
    mode, cur = None, None
    for ln in lines:
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
                ent = [(int(a), float(b)) for a, b in re.findall(r"\((\d+), ([-0-9.e+]+)\)", m.group(2))]
                cur.append([int(m.group(1)), ent])
        elif mode in ("B", "X"):
            try:
                g["b" if mode == "B" else "x"].append(float(ln))
            except ValueError:
                pass
    return g


def main():
    out = {"inputs": {}, "goldens": {}}
    for f in ("tridiag.inp", "identity.inp", "B.inp"):
        out["inputs"][f] = open(os.path.join(REF, f)).read()
    for p in sorted(glob.glob(os.path.join(REF, "*.ref"))):
        name = os.path.basename(p)[:-4]
        m = re.match(r"(\w+)-pc=(\w+)-metis=(\w+)(?:-opt=(\w+))?$", name)
        g = parse(p)
        g.update(input=m.group(1), pc=m.group(2), metis=m.group(3), opt=m.group(4) or "")
        out["goldens"][name] = g
    json.dump(out, open(OUT, "w"), indent=0, sort_keys=True)
    print("wrote", OUT, len(out["goldens"]), "goldens")


if __name__ == "__main__":
    main()
