"""Multi-GPU parity (needs >= 2 GPUs: run with `gpurun --gpus 2`): the distributed path (one process per GPU, NCCL halo /
allreduce / coarse gather inside the library) against the single-GPU path on the same box partition."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, lvl, ksp, q, partition="box"):
    try:
        import faulthandler
        faulthandler.dump_traceback_later(150, exit=True)  # a rank stuck in a collective shows where, then dies
        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch
        import torch.distributed as tdist
        import geneo4petsc_b200 as g
        from geneo4petsc_b200 import dist
        torch.cuda.set_device(rank)
        tdist.init_process_group("gloo", rank=rank, world_size=world)  # rendezvous only; the data path is the library's NCCL
        kind, args = "laplacian", "--dim 3 --size 32 --inpEps 0.0001 --kappa 2. lin"
        K, grid, sub_rank = dist.box_grid(world)
        nparts = len(sub_rank)
        if partition == "box":
            lo, hi = dist.keep_region(32, K, grid, rank)
            prob = g.Problem()
            edge = dist.generate_boxed(prob, kind, args, K, lo, hi)
            dist.decompose_owned(prob, nparts, sub_rank, rank, True, 0)
        else:  # the reference's METIS dual partition (src/geneo4PETSc.cpp:381-445), parts grouped onto the GPUs
            nparts, edge = 6, 32
            prob = g.Problem().generate(kind, args)
            sub_rank = dist.metis_problem(prob, nparts, world, rank)
        lay = dist.Layout(prob, rank, world, sub_rank)
        lay.exchange_requests(tdist)
        uid = dist.nccl_unique_id(tdist, rank)
        # tight eigen tolerance: with the default 1e-4 residual span(Z) is only reproducible to ~1e-5 between two runs that
        # differ in summation order (GenEO-2 eigenvalues cluster at the thresholds), far above the 1e-6 asked below
        opts = ["-geneo_lvl", lvl, "-geneo_tau", "0.2", "-els2_eps_tol", "1e-9"]
        pc = g.GeneoPC(opts)
        dist.setup_dist(pc, prob, lay, uid)
        n_own, n_loc = dist.local_sizes(pc)
        n = edge ** 3
        rng = np.random.default_rng(3)
        xg = rng.standard_normal(n)
        x = torch.zeros(n_loc, dtype=torch.float64, device="cuda")
        x[:n_own] = torch.from_numpy(xg[lay.owned]).cuda()
        y = torch.zeros_like(x)
        pc.mult_device(x.data_ptr(), y.data_ptr())
        torch.cuda.synchronize()
        ax = y[:n_own].cpu().numpy()
        pc.apply_device(x.data_ptr(), y.data_ptr())
        torch.cuda.synchronize()
        mx = y[:n_own].cpu().numpy()
        b = torch.zeros_like(x)
        ones = torch.zeros_like(x)
        ones[:n_own] = torch.from_numpy(lay.owned.astype(np.float64) + 1.0).cuda()
        pc.mult_device(ones.data_ptr(), b.data_ptr())
        sol = torch.zeros_like(x)
        r = pc.ksp_solve_device(b.data_ptr(), sol.data_ptr(), ksp=ksp, rtol=1e-8, atol=1e-50)
        torch.cuda.synchronize()
        info = pc.info()
        out = dict(rank=rank, owned=lay.owned, ax=ax, mx=mx, sol=sol[:n_own].cpu().numpy(), its=r["its"], reason=r["reason"],
                   nE=info["nE"], realDimE=info["realDimE"])
        gathered = [None] * world
        tdist.all_gather_object(gathered, out)
        if rank == 0:
            # single-GPU reference on the same partition
            if partition == "box":
                ref = g.Problem()
                dist.generate_boxed(ref, kind, args, K)
                eptr, eidx, _ = ref.mesh()
                first = eidx[eptr[:-1]]
                i, j, l = first % edge, (first // edge) % edge, first // (edge * edge)
                ep = ((i * K[0]) // edge + K[0] * ((j * K[1]) // edge + K[1] * ((l * K[2]) // edge))).astype(np.int32)
                ref.decompose(nparts, True, 0, elem_part=ep)
            else:
                ref = g.Problem().generate(kind, args)
                ref.decompose(nparts, True, 0)
                assert np.array_equal(ref.partition()[0], prob.partition()[0])  # identical subdomain partition
            pc1 = g.GeneoPC(opts).setup(ref)
            ax1, mx1 = pc1.mult(xg), pc1.apply(xg)
            r1 = pc1.ksp_solve(pc1.make_rhs(), ksp=ksp, rtol=1e-8, atol=1e-50)
            ax_d, mx_d, sol_d = np.zeros(n), np.zeros(n), np.zeros(n)
            for o in gathered:
                ax_d[o["owned"]] = o["ax"]; mx_d[o["owned"]] = o["mx"]; sol_d[o["owned"]] = o["sol"]
            np.testing.assert_allclose(ax_d, ax1, rtol=1e-11, atol=1e-11)
            assert pc1.info()["nE"] == gathered[0]["nE"] == gathered[1]["nE"]
            # GenEO-2 keeps eigenvectors on both sides of clustered thresholds (tau_loc, gamma_loc): which combination of a
            # nearly degenerate cluster is kept depends on the summation order, so M^-1 x is only reproducible to ~1e-5
            # between two builds (measured 1.2e-6 .. 1.2e-5 over three 2-GPU runs; same tolerance as the single-GPU
            # oracle comparison of GenEO-2); the Krylov SOLUTION below is compared to 1e-6 in every mode
            tol = 1e-4 if lvl.endswith("2") else 1e-6
            assert np.linalg.norm(mx_d - mx1) <= tol * np.linalg.norm(mx1), np.linalg.norm(mx_d - mx1) / np.linalg.norm(mx1)
            assert all(o["reason"] > 0 for o in gathered) and r1["reason"] > 0
            assert abs(gathered[0]["its"] - r1["its"]) <= 1, (gathered[0]["its"], r1["its"])
            assert np.linalg.norm(sol_d - r1["x"]) <= 1e-6 * np.linalg.norm(r1["x"])
        tdist.barrier()
        del pc
        tdist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "fail: %s\n%s" % (e, traceback.format_exc())))


@pytest.mark.parametrize("lvl,ksp,partition", [("ASM,1", "cg", "box"), ("ASM,H1", "gmres", "box"), ("SORAS,2", "gmres", "box"),
                                               ("ASM,1", "cg", "metis"), ("SORAS,2", "gmres", "metis")])
def test_two_gpu_solve_matches_single_gpu(lvl, ksp, partition):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 500) + {"ASM,1": 0, "ASM,H1": 1, "SORAS,2": 2}[lvl] + (5 if partition == "metis" else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lvl, ksp, q, partition)) for r in range(2)]
    for p in procs:
        p.start()
    import queue
    res = []
    try:
        for _ in procs:
            res.append(q.get(timeout=240))
            assert res[-1][1] == "ok", res[-1]  # report the first failure at once: the other rank is then stuck in a collective
    except queue.Empty:
        pytest.fail("a rank did not answer; got %r" % (res,))
    finally:
        for p in procs:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()
