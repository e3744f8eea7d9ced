"""Multi-GPU host logic on CPU: two gloo ranks build the same partition, assemble only their subdomains, exchange halo
requests and emulate the distributed operator in numpy.  (The NCCL data path itself needs GPUs: tests/test_gpu_dist.py.)"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _global_operator(g, kind, args, nparts, dual=True, overlap=0, part=None):
    p = g.Problem().generate(kind, args)
    if part is None:
        p.decompose(nparts, dual, overlap)
    else:
        p.decompose(nparts, dual, overlap, elem_part=part)
    n = p.sizes()["nb_node"]
    import scipy.sparse as sp
    a = sp.csr_matrix((n, n))
    for s in range(nparts):
        nodes, _ = p.sub_nodes(s)
        r = sp.csr_matrix((np.ones(len(nodes)), (np.arange(len(nodes)), nodes)), shape=(len(nodes), n))
        a = a + r.T @ p.sub_matrix(s, 0) @ r
    return p, a.tocsr()


def _worker(rank, world, port, mode, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch.distributed as tdist
        import geneo4petsc_b200 as g
        from geneo4petsc_b200 import dist
        tdist.init_process_group("gloo", rank=rank, world_size=world)
        kind, args = "laplacian", "--dim 3 --size 10 --inpEps 0.0001 --kappa 2. lin"
        if mode == "metis":
            nparts = 4
            full, a_glob = _global_operator(g, kind, args, nparts)
            ep, _ = full.partition()
            sub_rank = np.array([0, 0, 1, 1], dtype=np.int32)
            prob = g.Problem().generate(kind, args)
            dist.decompose_owned(prob, nparts, sub_rank, rank, True, 0, elem_part=ep)
        elif mode == "metis_grouped":  # METIS k-way, parts grouped onto the ranks along the heaviest interfaces
            nparts = 6
            full, a_glob = _global_operator(g, kind, args, nparts)
            prob = g.Problem().generate(kind, args)
            sub_rank = dist.metis_problem(prob, nparts, world, rank)
            assert np.array_equal(prob.partition()[0], full.partition()[0])  # METIS is deterministic: the ranks agree
            assert sorted(np.bincount(sub_rank, minlength=world).tolist()) == [3, 3]
            ep, ei, _ = prob.mesh()
            w = dist.part_adjacency(ep, ei, prob.partition()[0], nparts)
            cut = sum(w[a, b] for a in range(nparts) for b in range(nparts) if sub_rank[a] != sub_rank[b]) // 2
            worst = max(sum(w[a, b] for a in range(nparts) for b in range(nparts) if sr[a] != sr[b]) // 2
                        for sr in ([0, 1, 0, 1, 0, 1], [0, 0, 0, 1, 1, 1], [1, 0, 0, 1, 1, 0]))
            assert cut <= worst  # the greedy grouping is no worse than blind assignments
        else:  # box partition, each rank generates only its sub-mesh
            K, grid, sub_rank = dist.box_grid(world)
            nparts = len(sub_rank)
            ref = g.Problem()
            edge = dist.generate_boxed(ref, kind, args, K)
            ep_full = np.zeros(ref.sizes()["nb_elem"], dtype=np.int32)
            # partition of the FULL mesh through the same rule (first node of the element)
            eptr, eidx, _ = ref.mesh()
            first = eidx[eptr[:-1]]
            i, j, l = first % edge, (first // edge) % edge, first // (edge * edge)
            ep_full = ((i * K[0]) // edge + K[0] * ((j * K[1]) // edge + K[1] * ((l * K[2]) // edge))).astype(np.int32)
            _, a_glob = _global_operator(g, kind, args, nparts, part=ep_full)
            lo, hi = dist.keep_region(edge, K, grid, rank)
            prob = g.Problem()
            dist.generate_boxed(prob, kind, args, K, lo, hi)
            assert prob.sizes()["nb_elem"] < ref.sizes()["nb_elem"]
            dist.decompose_owned(prob, nparts, sub_rank, rank, True, 0)
        lay = dist.Layout(prob, rank, world, sub_rank)
        asked = lay.exchange_requests(tdist)  # set_send() validates that every requested node is owned here
        n = a_glob.shape[0]
        owned_all = [None] * world
        tdist.all_gather_object(owned_all, lay.owned.tolist())
        allo = np.sort(np.concatenate([np.array(o, dtype=np.int64) for o in owned_all]))
        assert np.array_equal(allo, np.arange(n)), "owned sets must partition the nodes"
        # subdomain matrices of the local subdomains equal the single-process ones
        # distributed SpMV emulation: ghosts filled from the owners' values
        x = np.random.default_rng(7).standard_normal(n)
        xl = np.concatenate([x[lay.owned], x[lay.ghost]])
        y = lay.matrix() @ xl
        np.testing.assert_allclose(y, (a_glob @ x)[lay.owned], rtol=1e-12, atol=1e-12)
        # halo lists are mutually consistent: what q asked of me == my send list, element by element
        for qq in range(world):
            if qq != rank:
                assert np.all(np.isin(asked[qq], lay.owned))
        tdist.barrier()
        tdist.destroy_process_group()
        q.put((rank, "ok", lay.n_own, lay.n_ghost))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "fail: %s\n%s" % (e, traceback.format_exc()), 0, 0))


@pytest.mark.parametrize("mode", ["metis", "box", "metis_grouped"])
def test_two_rank_layout_gloo(mode):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + {"metis": 0, "box": 1000, "metis_grouped": 2000}[mode]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] == "ok", r
    assert sum(r[2] for r in res) == 1000
    assert all(r[3] > 0 for r in res)
