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

def factor3(n):
    """n = a * b * c with a >= b >= c as close to each other as possible (box grids of ranks and of subdomains per rank)."""
    best = None
    for c in range(1, int(round(n ** (1.0 / 3.0))) + 2):
        if n % c:
            continue
        m = n // c
        for b in range(c, int(m ** 0.5) + 1):
            if m % b:
                continue
            a = m // b
            if best is None or a - c < best[0] - best[2]:
                best = (a, b, c)
    return best if best else (n, 1, 1)


def box_dims(nparts):
    """boxes per axis of a single-process box partition into nparts subdomains"""
    return factor3(nparts)


def box_grid(world, subs_per_gpu=8):
    """Global box grid K, rank grid g, per-rank block of boxes blk, and the rank owning every box: each rank gets a
    compact blk[0] x blk[1] x blk[2] block of boxes (2x2x2 for the default 8 subdomains per GPU)."""
    g = factor3(world)
    blk = factor3(subs_per_gpu)
    K = (g[0] * blk[0], g[1] * blk[1], g[2] * blk[2])
    sub_rank = np.zeros(K[0] * K[1] * K[2], dtype=np.int32)
    for b3 in range(K[2]):
        for b2 in range(K[1]):
            for b1 in range(K[0]):
                r = (b1 // blk[0]) + g[0] * ((b2 // blk[1]) + g[1] * (b3 // blk[2]))
                sub_rank[b1 + K[0] * (b2 + K[1] * b3)] = r
    return K, g, sub_rank


def keep_region(edge, K, g, rank, subs_per_gpu=8):
    """Node-coordinate region [lo, hi) a rank must hold: its boxes, one node beyond their upper faces (elements hang on
    their lower node) and one more ring for the couplings of A_dir / owned rows."""
    blk = factor3(subs_per_gpu)
    r = (rank % g[0], (rank // g[0]) % g[1], rank // (g[0] * g[1]))
    lo, hi = [], []
    for a in range(3):
        b0, b1 = blk[a] * r[a], blk[a] * (r[a] + 1)  # boxes [b0, b1) along axis a
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
# General partitions on several GPUs (METIS k-way or any caller-supplied partition): which rank holds which subdomain?
# The reference has one MPI rank per subdomain (src/geneo4PETSc.cpp:604); here a rank = a GPU = several subdomains, so the
# subdomains are GROUPED: equal counts per rank, grown greedily along the heaviest interfaces so that most neighbours of a
# subdomain sit on the same GPU (halo traffic = the cut between the groups).
# ---------------------------------------------------------------------------------------------------------------------
def part_adjacency(elem_ptr, elem_idx, elem_part, nb_part):
    """W[p, q] = number of mesh nodes shared by the elements of parts p and q (p != q); dual (element) partitions."""
    import scipy.sparse as sp
    elem_ptr = np.asarray(elem_ptr, dtype=np.int64)
    elem_idx = np.asarray(elem_idx, dtype=np.int64)
    elem_of = np.repeat(np.arange(len(elem_ptr) - 1), np.diff(elem_ptr))
    nn = int(elem_idx.max()) + 1
    inc = sp.csr_matrix((np.ones(len(elem_idx), dtype=np.int8), (elem_idx, np.asarray(elem_part, dtype=np.int64)[elem_of])),
                        shape=(nn, nb_part))
    inc.data[:] = 1
    inc.sum_duplicates()
    inc.data[:] = 1
    w = (inc.T @ inc).toarray().astype(np.int64)
    np.fill_diagonal(w, 0)
    return w


def assign_ranks(w, world):
    """sub_rank[p] from the interface weights w (nb_part x nb_part, symmetric): groups of ceil/floor(nb_part / world)
    subdomains, each grown from the unassigned subdomain with the least outside contact by repeatedly taking the unassigned
    subdomain most strongly tied to the group.  Deterministic (ties -> lowest index): every rank computes the same map."""
    w = np.asarray(w, dtype=np.float64)
    P = w.shape[0]
    sub_rank = -np.ones(P, dtype=np.int32)
    sizes = [P // world + (1 if r < P % world else 0) for r in range(world)]
    for r in range(world):
        free = np.flatnonzero(sub_rank < 0)
        if len(free) == 0:
            break
        contact = w[np.ix_(free, free)].sum(axis=1)
        seed = free[int(np.argmin(contact))]
        sub_rank[seed] = r
        for _ in range(sizes[r] - 1):
            free = np.flatnonzero(sub_rank < 0)
            if len(free) == 0:
                break
            tie = w[np.ix_(free, np.flatnonzero(sub_rank == r))].sum(axis=1)
            sub_rank[free[int(np.argmax(tie))]] = r
    sub_rank[sub_rank < 0] = world - 1
    return sub_rank


def metis_problem(problem, nb_part, world, rank, dual=True, overlap=0):
    """Every rank partitions the SAME (already generated / read) mesh with METIS -- deterministic, so the ranks agree without
    talking --, groups the parts onto the ranks (assign_ranks) and assembles only its own subdomains.  Returns sub_rank."""
    probe = api.Problem()
    ep, ei, em = problem.mesh()
    s = problem.sizes()
    probe.set_mesh(s["nb_node"], ep, ei, em)
    from .api import lib as _lib
    # partition only: METIS through the library's own entry (the decomposition of ALL subdomains is not needed for it)
    epart = np.zeros(s["nb_elem"], dtype=np.int32)
    npart = np.zeros(s["nb_node"], dtype=np.int32)
    _chk(_lib.geneo_problem_partition(probe.h, C.c_int(nb_part), C.c_int(1 if dual else 0), _p(epart, _i32p), _p(npart, _i32p)))
    del probe
    if dual:
        w = part_adjacency(ep, ei, epart, nb_part)
    else:
        w = part_adjacency(ep, ei, npart[ei[np.asarray(ep[:-1], dtype=np.int64)]], nb_part)
    sub_rank = assign_ranks(w, world)
    decompose_owned(problem, nb_part, sub_rank, rank, dual, overlap, elem_part=epart if dual else None, node_part=None if dual else npart)
    return sub_rank
