"""Self-consistency KATs of the oracle and generator restatement vs the reference's own generators."""
import os

import numpy as np
import pytest
import scipy.linalg as sla

from oracle import geneo_oracle as go

HAVE_REF = os.path.exists(os.path.join(os.path.dirname(go.__file__), "_ref", "libgenlaplacian.so"))


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference once)")
@pytest.mark.parametrize("kind,args,kw", [
    ("laplacian", "--dim 3 --size 10 --kappa 2. lin --inpEps 0.0001", dict(dim=3, size=10, inp_eps=1e-4, kappa_max=2.0, interp="lin")),
    ("laplacian", "--dim 2 --size 12", dict(dim=2, size=12)),
    ("laplacian", "--dim 1 --size 9 --weakScaling 2", dict(dim=1, size=9, weak=2)),
    ("laplacian", "--dim 3 --size 5 --weakScaling 2 --kappa 7. quad", dict(dim=3, size=5, weak=2, kappa_max=7.0, interp="quad")),
    ("heat", "--dim 2 --size 7 --kappa 100. minmax --lbd 1. --dt 0.1", dict(dim=2, size=7, kappa_max=100.0, interp="minmax", heat=True)),
    ("heat", "--dim 3 --size 6 --weakScaling 2 --kappa 3. quad --lbd 2. --dt 0.5 --inpEps 0.01",
     dict(dim=3, size=6, inp_eps=0.01, kappa_max=3.0, interp="quad", weak=2, heat=True, lbd=2.0, dt=0.5)),
])
def test_generator_restatement_is_bit_exact(kind, args, kw):
    a, b = go.ref_generator(kind, args), go.gen_grid(**kw)
    assert a.nb_node == b.nb_node and a.nb_elem == b.nb_elem
    assert np.array_equal(a.elem_ptr, b.elem_ptr) and np.array_equal(a.elem_idx, b.elem_idx)
    assert np.array_equal(a.mat_val, b.mat_val)


@pytest.fixture(scope="module")
def case3d():
    mesh = go.gen_grid(3, 10, 1e-4, 2.0, "lin")
    return mesh


def test_operator_is_sum_of_weighted_neumann(case3d):
    mesh = case3d
    for dual, ov in [(True, 0), (False, 0), (True, 1)]:
        part = go.metis_partition(mesh, 3, dual)
        dec = go.decompose(mesh, 3, part[0], part[1], dual, ov)
        a = go.assemble_global(mesh.nb_node, dec, [go.local_neumann(mesh, dec, p) for p in range(3)])
        # the assembled operator does not depend on the decomposition: compare with a 1-domain assembly
        p1 = go.metis_partition(mesh, 1, True)
        d1 = go.decompose(mesh, 1, p1[0], p1[1], True, 0)
        a1 = go.local_neumann(mesh, d1, 0)
        assert abs(a - a1).max() < 1e-12


def test_inertia_equals_eigen_count_and_coarse_identities(case3d):
    mesh = case3d
    rep = go.run_case(mesh, 4, go.GenEOOptions(tau=0.3), ksp="cg", rtol=1e-8, atol=1e-50)
    pc = rep.pc
    for s in pc.sub:
        dd = np.diag(s.d)
        w = sla.eigvalsh(s.a_neu.toarray(), dd @ s.a_dir.toarray() @ dd)
        assert s.estim == int(np.sum(w < 0.3))
        assert len([v for v in s.eigvals if v > 0]) == s.estim
        np.testing.assert_allclose(sorted(s.eigvals)[: s.estim], w[: s.estim], rtol=1e-8)
    z = pc.zmat.toarray()
    a = rep.a.toarray()
    q = np.stack([pc.apply_q(a[:, j]) for j in range(a.shape[1])], axis=1)  # Q A
    np.testing.assert_allclose(q @ z, z, atol=1e-8)  # Q A Z = Z  <=> (I-P) Z = 0
    assert rep.ksp.converged and rep.true_rel_res < 1e-6


@pytest.mark.parametrize("lvl1,lvl2", [("ASM", "1"), ("ASM", "H1"), ("SRAS", "1"), ("SORAS", "2")])
def test_symmetric_variants_are_symmetric(case3d, lvl1, lvl2):
    mesh = case3d
    rep = go.run_case(mesh, 3, go.GenEOOptions(lvl1=lvl1, lvl2=lvl2), ksp="cg", rtol=1e-6)
    rng = np.random.default_rng(0)
    u, v = rng.standard_normal(mesh.nb_node), rng.standard_normal(mesh.nb_node)
    assert abs(u @ rep.pc.apply(v) - v @ rep.pc.apply(u)) < 1e-8 * abs(u @ rep.pc.apply(v))


def test_cg_and_gmres_agree(case3d):
    mesh = case3d
    r1 = go.run_case(mesh, 4, go.GenEOOptions(), ksp="cg", rtol=1e-10)
    r2 = go.run_case(mesh, 4, go.GenEOOptions(), ksp="gmres", rtol=1e-10, restart=1000)
    np.testing.assert_allclose(r1.ksp.x, np.arange(1, mesh.nb_node + 1.0), rtol=1e-6)
    np.testing.assert_allclose(r2.ksp.x, np.arange(1, mesh.nb_node + 1.0), rtol=1e-6)
    assert abs(r1.ksp.its - r2.ksp.its) <= 2
