"""Multi-GPU host plumbing: one process per GPU (torchrun), torch.distributed for the rendezvous only.

The hot path itself stays inside libgeneob200.so: halo exchanges, Krylov allreduces and the coarse gather are NCCL
calls issued by the C++ side (csrc/comm.cu).  This module (a) builds the same partition on every rank, (b) lets every
rank assemble only its own subdomains, (c) exchanges the halo requests once, (d) broadcasts rank 0's NCCL unique id.

Replaces what `mpirun -n P` + Boost.MPI send/recv of the domains do in the reference (src/geneo4PETSc.cpp:496-569, 604):
there, one rank = one subdomain; here one rank = one GPU = several subdomains.
"""
import ctypes as C
import json
import os
import time

import numpy as np

from . import api
from .api import lib, _chk, _p, _i32p, _i64p, _f64p

RANK_GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def box_grid(world, subs_per_gpu=8):
    """Global box grid and the rank owning every box: each rank gets a compact 2x2x2 block of boxes."""
    if subs_per_gpu != 8 or world not in RANK_GRIDS:
        raise ValueError("box partition: 8 subdomains per GPU and 1/2/4/8 GPUs")
    g = RANK_GRIDS[world]
    K = (2 * g[0], 2 * g[1], 2 * g[2])
    sub_rank = np.zeros(K[0] * K[1] * K[2], dtype=np.int32)
    for b3 in range(K[2]):
        for b2 in range(K[1]):
            for b1 in range(K[0]):
                r = (b1 // 2) + g[0] * ((b2 // 2) + g[1] * (b3 // 2))
                sub_rank[b1 + K[0] * (b2 + K[1] * b3)] = r
    return K, g, sub_rank


def keep_region(edge, K, g, rank):
    """Node-coordinate region [lo, hi) a rank must hold: its boxes, one node beyond their upper faces (elements hang on
    their lower node) and one more ring for the couplings of A_dir / owned rows."""
    r = (rank % g[0], (rank // g[0]) % g[1], rank // (g[0] * g[1]))
    lo, hi = [], []
    for a in range(3):
        b0, b1 = 2 * r[a], 2 * r[a] + 2  # boxes [b0, b1) along axis a
        # box of coordinate i is (i*K)//edge  ->  first coordinate of box b is ceil(b*edge/K)
        first = -(-b0 * edge // K[a])
        last = -(-b1 * edge // K[a])  # first coordinate of the next block (exclusive)
        lo.append(max(0, first - 1))
        hi.append(min(edge, last + 2))
    return np.array(lo, dtype=np.int32), np.array(hi, dtype=np.int32)


class Layout:
    def __init__(self, problem, rank, world, sub_rank):
        self.h = C.c_void_p()
        self.rank, self.world = rank, world
        sr = np.ascontiguousarray(sub_rank, dtype=np.int32)
        _chk(lib.geneo_layout_create(problem.h, C.c_int(rank), C.c_int(world), _p(sr, _i32p), C.byref(self.h)))
        s = np.zeros(4, dtype=np.int64)
        _chk(lib.geneo_layout_sizes(self.h, _p(s, _i64p)))
        self.n_own, self.n_ghost, self.nnz = int(s[0]), int(s[1]), int(s[2])
        self.owned = np.zeros(self.n_own, dtype=np.int32)
        self.ghost = np.zeros(max(1, self.n_ghost), dtype=np.int32)
        self.ghost_ptr = np.zeros(world + 1, dtype=np.int64)
        _chk(lib.geneo_layout_get(self.h, _p(self.owned, _i32p), _p(self.ghost, _i32p), _p(self.ghost_ptr, _i64p)))
        self.ghost = self.ghost[: self.n_ghost]

    def __del__(self):
        if getattr(self, "h", None) and self.h:
            lib.geneo_layout_destroy(self.h)
            self.h = None

    def matrix(self):
        import scipy.sparse as sp
        ptr, idx, val = np.zeros(self.n_own + 1, dtype=np.int64), np.zeros(self.nnz, dtype=np.int32), np.zeros(self.nnz)
        _chk(lib.geneo_layout_matrix(self.h, _p(ptr, _i64p), _p(idx, _i32p), _p(val, _f64p)))
        return sp.csr_matrix((val, idx, ptr), shape=(self.n_own, self.n_own + self.n_ghost))

    def requests(self):
        """per peer: the global ids this rank needs from it (its ghosts owned by that peer)"""
        return [self.ghost[self.ghost_ptr[q]: self.ghost_ptr[q + 1]].copy() for q in range(self.world)]

    def set_send(self, peer, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        _chk(lib.geneo_layout_set_send(self.h, C.c_int(peer), _p(ids, _i32p), C.c_int64(len(ids))))

    def exchange_requests(self, tdist):
        """all-to-all of the halo requests through torch.distributed (gloo or nccl); returns what each peer asked for."""
        mine = [r.tolist() for r in self.requests()]
        everyone = [None] * self.world
        tdist.all_gather_object(everyone, mine)
        asked = [np.array(everyone[q][self.rank], dtype=np.int32) for q in range(self.world)]
        for q in range(self.world):
            if q != self.rank:
                self.set_send(q, asked[q])
        return asked


def decompose_owned(problem, nb_part, sub_rank, rank, dual=True, overlap=0, elem_part=None, node_part=None):
    ep = None if elem_part is None else np.ascontiguousarray(elem_part, dtype=np.int32)
    npart = None if node_part is None else np.ascontiguousarray(node_part, dtype=np.int32)
    sr = np.ascontiguousarray(sub_rank, dtype=np.int32)
    _chk(lib.geneo_problem_decompose_owned(problem.h, C.c_int(nb_part), C.c_int(1 if dual else 0), C.c_int(overlap), _p(ep, _i32p),
                                           _p(npart, _i32p), _p(sr, _i32p), C.c_int(rank)))
    problem._sizes_hint = nb_part
    return problem


def generate_boxed(problem, kind, args, K, keep_lo=None, keep_hi=None):
    k = np.ascontiguousarray(K, dtype=np.int32)
    edge = C.c_int32()
    lo = None if keep_lo is None else np.ascontiguousarray(keep_lo, dtype=np.int32)
    hi = None if keep_hi is None else np.ascontiguousarray(keep_hi, dtype=np.int32)
    _chk(lib.geneo_problem_generate_boxed(problem.h, kind.encode(), args.encode(), _p(k, _i32p), _p(lo, _i32p), _p(hi, _i32p), C.byref(edge)))
    return edge.value


def nccl_unique_id(tdist, rank):
    """rank 0 creates the id, everyone receives it (128 bytes)."""
    buf = (C.c_char * 128)()
    if rank == 0:
        _chk(lib.geneo_nccl_unique_id(buf))
    box = [bytes(buf)]
    tdist.broadcast_object_list(box, src=0)
    return box[0]


def setup_dist(pc, problem, layout, uid):
    pc.problem, pc.layout = problem, layout  # borrowed by the library: keep them alive
    _chk(lib.geneo_pc_setup_dist(pc.h, problem.h, layout.h, C.c_char_p(uid) if uid is not None else None))
    return pc


def local_sizes(pc):
    s = np.zeros(2, dtype=np.int64)
    _chk(lib.geneo_pc_local_sizes(pc.h, _p(s, _i64p)))
    return int(s[0]), int(s[1])


def allreduce_sum(pc, values):
    v = np.ascontiguousarray(values, dtype=np.float64).copy()
    _chk(lib.geneo_allreduce_sum(pc.h, _p(v, _f64p), C.c_int(len(v))))
    return v


# ---------------------------------------------------------------------------------------------------------------------
# bench.py, N > 1  (weak scaling: 8 subdomains and size^3 DOFs per GPU, box partition, each rank generates its sub-mesh)
# ---------------------------------------------------------------------------------------------------------------------
def build_rank_problem(kind, gen_args, rank, world, subs_per_gpu=8, overlap=0):
    import torch.distributed as tdist
    K, g, sub_rank = box_grid(world, subs_per_gpu)
    prob = api.Problem()
    edge = generate_boxed(prob, kind, gen_args, K)  # cheap probe of the edge? no: generate once below with the region
    return prob, edge, K, g, sub_rank, tdist


def run_bench(a, rank, world, local, METRIC, UNIT, config, ClockSampler):
    import torch
    import torch.distributed as tdist
    tdist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    from bench import gen_args, weak_edge
    edge = weak_edge(a.size, world)
    K, g, sub_rank = box_grid(world, a.subs_per_gpu)
    nb_part = len(sub_rank)
    t0 = time.time()
    lo, hi = keep_region(edge, K, g, rank)
    prob = api.Problem()
    e2 = generate_boxed(prob, a.kind, gen_args(a, edge), K, lo, hi)
    assert e2 == edge, (e2, edge)
    t1 = time.time()
    decompose_owned(prob, nb_part, sub_rank, rank, True, 0)
    layout = Layout(prob, rank, world, sub_rank)
    layout.exchange_requests(tdist)
    uid = nccl_unique_id(tdist, rank)
    t2 = time.time()
    n = edge ** 3
    opts = ["-geneo_lvl", a.lvl, "-geneo_tau", a.tau, "-geneo_kernel_timing"]
    pc = api.GeneoPC(opts)
    setup_dist(pc, prob, layout, uid)
    tm_cold = pc.timers()
    st = pc.stats()
    n_own, n_loc = local_sizes(pc)
    x = torch.zeros(n_loc, dtype=torch.float64, device="cuda")
    b = torch.zeros(n_loc, dtype=torch.float64, device="cuda")
    ones = torch.zeros(n_loc, dtype=torch.float64, device="cuda")
    ones[:n_own] = torch.from_numpy(layout.owned.astype(np.float64) + 1.0).cuda()
    pc.mult_device(ones.data_ptr(), b.data_ptr())  # b = A (1..N)
    torch.cuda.synchronize()

    def step():
        pc.refactor()
        return pc.ksp_solve_device(b.data_ptr(), x.data_ptr(), ksp=a.ksp, rtol=a.rtol, atol=1e-50, restart=30)

    for _ in range(a.warmup):
        r = step()
    pc.kernel_time()
    c0 = api.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        tdist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.steps):
            r = step()
        e1.record()
        torch.cuda.synchronize()
        tdist.barrier()
    ms = e0.elapsed_time(e1) / a.steps
    c1 = api.counters()
    kms, klaunch = pc.kernel_time()
    err_loc = float((x[:n_own] - ones[:n_own]).abs().max() / n)
    # e2e: create + setup (host symbolic, uploads, numeric) + solve with host buffers, through the C ABI
    bh = b.cpu().numpy()
    tdist.barrier()
    ta = time.perf_counter()
    k0 = api.counters()
    pc2 = api.GeneoPC(["-geneo_lvl", a.lvl, "-geneo_tau", a.tau])
    uid2 = nccl_unique_id(tdist, rank)
    setup_dist(pc2, prob, layout, uid2)
    r2 = pc2.ksp_solve(bh, ksp=a.ksp, rtol=a.rtol, atol=1e-50, restart=30)
    torch.cuda.synchronize()
    tb = time.perf_counter()
    k1 = api.counters()
    tm2 = pc2.timers()
    del pc2
    # max over ranks (device time of the steps, wall time of e2e), sums of the per-rank statistics
    red = torch.tensor([ms, tb - ta, err_loc, kms / max(1, klaunch)], dtype=torch.float64, device="cuda")
    tdist.all_reduce(red, op=tdist.ReduceOp.MAX)
    sums = torch.tensor([st["trisolve_bytes"], st["factor_bytes"], st["factor_flops"], float(k1["h2d"] - k0["h2d"]),
                         float(k1["d2h"] - k0["d2h"]), float(c1["launches"] - c0["launches"])], dtype=torch.float64, device="cuda")
    tdist.all_reduce(sums, op=tdist.ReduceOp.SUM)
    ms, e2e_s, err, kavg = [float(v) for v in red.cpu()]
    tri_b, fac_b, fac_f, h2d, d2h, launches = [float(v) for v in sums.cpu()]
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        ach = st["trisolve_bytes"] / kavg / 1e6 if kavg > 0 else 0.0  # rank 0's kernel against one GPU's HBM
        out = {
            "metric": METRIC, "value": n / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config, partition="box %dx%dx%d (METIS on the global mesh does not fit one rank at N>1)" % K),
            "clocks": clk.summary(),
            "e2e": {"value": n / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "seconds": e2e_s,
                    "symbolic_s": tm2["symbolic"], "upload_s": tm2["upload"], "numeric_s": tm2["numeric"]},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "k_solve_ring<1> (rank 0)", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": None, "algorithmic_bytes_per_launch": st["trisolve_bytes"],
                         "launches_timed": klaunch, "avg_launch_ms": kavg,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"},
            "detail": {"n_dof": n, "iterations": r["its"], "reason": r["reason_name"], "rnorm": r["rnorm"], "max_rel_err_vs_1..N": err,
                       "dimE": pc.info()["nE"], "rank0_local": {"n_own": n_own, "n_ghost": n_loc - n_own},
                       "cold_setup_rank0": {k: tm_cold[k] for k in ("symbolic", "upload", "numeric", "operator", "setup")},
                       "gen_s": t1 - t0, "decomp_layout_s": t2 - t1, "factor_bytes_total": fac_b, "factor_flops_total": fac_f,
                       "trisolve_bytes_total": tri_b},
        }
        print(json.dumps(out), flush=True)
    tdist.barrier()
    tdist.destroy_process_group()
