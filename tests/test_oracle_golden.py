"""Pin the oracle against the reference's own goldens (tst/dummy/*.ref, 84 files).

The goldens fix: the per-rank local Neumann matrices (=> METIS partition + overlap + 1/mult element weighting),
the RHS b = A (1..N), the converged solution, the counts of the first INFO line, the KSP line and the PC name string.
METIS nodal partitions from the CUDA-toolkit libmetis come out with the two labels swapped w.r.t. the goldens
 so local matrices are compared up to a relabelling of the subdomains.
"""
import itertools

import numpy as np
import pytest

from oracle import geneo_oracle as go
from tests._cases import dense_from_golden, golden_config


def _names(goldens):
    return sorted(goldens["goldens"].keys())


def _run(g, inputs):
    cfg = golden_config(g)
    eps = 1.0 if g["input"] == "tridiag" else 1e-4  # tst/dummy/dummy.sh:64
    mesh = go.read_input_file(str(inputs / (g["input"] + ".inp")), eps)
    b = go.read_rhs_file(str(inputs / "B.inp"), mesh.nb_node) if g["input"] == "identity" else None
    return mesh, b, cfg


def test_all_84_goldens_present(dummy_goldens):
    assert len(dummy_goldens["goldens"]) == 84


@pytest.mark.parametrize("idx", range(84))
def test_dummy_golden(idx, dummy_goldens, dummy_inputs):
    name = _names(dummy_goldens)[idx]
    g = dummy_goldens["goldens"][name]
    mesh, b, cfg = _run(g, dummy_inputs)
    dual = g["metis"] == "dual"
    if cfg is None:  # bjacobi: only the assembled operator, RHS and solution are in scope
        part = go.metis_partition(mesh, 2, dual)
        dec = go.decompose(mesh, 2, part[0], part[1], dual, 0)
        a = go.assemble_global(mesh.nb_node, dec, [go.local_neumann(mesh, dec, p) for p in range(2)]).toarray()
        np.testing.assert_allclose(a, dense_from_golden(g["mats"][0], 8), atol=1e-14)
        bb = b if b is not None else a @ np.arange(1, 9.0)
        np.testing.assert_allclose(bb, g["b"], atol=1e-12)
        np.testing.assert_allclose(np.linalg.solve(a, bb), g["x"], atol=1e-5)
        return
    opt = go.GenEOOptions(lvl1=cfg["lvl1"], lvl2=cfg["lvl2"], cut=10 if g["input"] == "tridiag" else -1,
                          offload=cfg["offload"])
    rep = go.run_case(mesh, 2, opt, dual=cfg["dual"], overlap=cfg["overlap"], ksp="gmres", rtol=1e-12, atol=1e-12, b=b)
    # local matrices (up to subdomain relabelling)
    mine = [m.toarray() for m in rep.pc.a_neu]
    gold = [dense_from_golden(rows, len(rows)) for rows in g["mats"]]
    ok = False
    for perm in itertools.permutations(range(2)):
        if all(mine[perm[i]].shape == gold[i].shape and np.allclose(mine[perm[i]], gold[i], atol=1e-14) for i in range(2)):
            ok = True
    assert ok, "local Neumann matrices differ from golden " + name
    np.testing.assert_allclose(rep.b, g["b"], atol=1e-12)
    assert rep.ksp.converged
    np.testing.assert_allclose(rep.ksp.x, g["x"], atol=2e-5)  # goldens print PETSc %g precision
    info0 = "INFO: nb DOFs %d, nb elements %d, nnz coefs %d, nb partitions %d, overlap %d, metis %s" % (
        mesh.nb_node, mesh.nb_elem, rep.nnz_loc, 2, cfg["overlap"], "dual" if cfg["dual"] else "nodal")
    assert info0 == g["info"][0]
    assert g["info"][2].startswith("INFO: %s pc" % opt.name())
    assert g["info"][3] == "INFO: solve - converged"
