#!/usr/bin/env python
"""bench.py -- the driver-facing benchmark of the GenEO hot path (contract: see DESIGN.md "Measurement").

Workload (BASELINE.json): 3-D 7-point Laplacian from the reference's own generator grammar
(`--dim 3 --size S --inpEps 0.0001`, tst/laplacian/laplacian.cpp:56-188), METIS-dual partition into 8 subdomains per
GPU, two-level GenEO (`-geneo_lvl ASM,1 -geneo_tau 0.1`), CG to rtol 1e-5 on the preconditioned residual.
Default S = 200 (8.0 M DOFs per GPU): configs[1] (256^3 on one GPU, 8 x 128^3 subdomains) needs 147 GB of FP64
factors + 18 GB transient factor + 27 GB update arenas = 192 GB > 180 GB HBM (symbolic analysis of a 128^3 subdomain,
DESIGN.md "Memory"), so the largest single-GPU configuration of BASELINE.json -- configs[4], 8 M DOFs per GPU,
the weak-scaling unit -- is the N=1 workload.

One "step" = one complete pass of the hot path over the problem resident in HBM:
    numeric GenEO setup (3 sparse LDL^T factorizations + 1 block-Lanczos eigen-solve per subdomain, Z, E = Z^T A Z,
    E^-1) followed by the preconditioned Krylov solve                       -> `value` = N_dof / step seconds.
`e2e` = the same metric through the C ABI with HOST buffers: geneo_pc_create + geneo_pc_setup (host symbolic analysis,
H2D upload of every matrix, numeric setup) + geneo_ksp_solve (host b in, host x out) -- what a PETSc caller times as
"solver set up" + "solver iterations" (src/geneo4PETSc.cpp:1363-1367, 1239-1242).

`--impl reference` times the CPU restatement of the reference (oracle/, scipy SuperLU + ARPACK; the reference itself
needs PETSc/SLEPc/MUMPS/MPI which this image does not have) on the host cores, on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "geneo_solve_throughput(setup+iter)"
UNIT = "DOF/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=200, help="grid edge per GPU (weak scaling: edge = cbrt(size^3 * gpus))")
    ap.add_argument("--kind", default="laplacian", choices=["laplacian", "heat"])
    ap.add_argument("--subs-per-gpu", type=int, default=8)
    ap.add_argument("--lvl", default="ASM,1")
    ap.add_argument("--tau", default="0.1")
    ap.add_argument("--ksp", default="cg")
    ap.add_argument("--rtol", type=float, default=1e-5)
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--cpu-size", type=int, default=48, help="grid edge of the bounded CPU sample (cpu_baseline / reference arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def gen_args(a, size):
    s = "--dim 3 --size %d --inpEps 0.0001" % size
    if a.kind == "heat":
        s += " --kappa 100. minmax --lbd 1. --dt 0.1"
    return s


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md, "clocks line")."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.th.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (oracle = test infrastructure; used here ONLY as the timed CPU baseline, never on the product path)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_sample(a, size, nparts):
    """One setup+solve of the CPU restatement on `size`^3 DOFs; returns (dofs, seconds, iterations, cores)."""
    from oracle import geneo_oracle as go
    kappa, interp = (100.0, "minmax") if a.kind == "heat" else (1.0, "")
    mesh = go.gen_grid(3, size, 1e-4, kappa, interp, heat=(a.kind == "heat"))
    l1, l2 = a.lvl.split(",")
    cores = max(1, min(nparts, os.cpu_count() or 1))  # one worker per subdomain = the reference's one MPI rank per subdomain
    rep = go.run_case(mesh, nparts, go.GenEOOptions(lvl1=l1, lvl2=l2, tau=float(a.tau)), ksp=a.ksp, rtol=a.rtol, workers=cores)
    return mesh.nb_node, rep.setup_s + rep.solve_s, rep.ksp.its, cores


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, its, n = [], 0, 0
    for i in range(a.warmup + a.steps):
        n, secs, its, cores = cpu_sample(a, a.cpu_size, a.subs_per_gpu)
        if i >= a.warmup:
            vals.append(secs)
    t = sum(vals) / len(vals)
    v = n / t
    sample = "%s %d^3 = %d DOFs, %d subdomains, %s, %s rtol %g (%d its): scipy SuperLU/ARPACK restatement" % (
        a.kind, a.cpu_size, n, a.subs_per_gpu, a.lvl, a.ksp, a.rtol, its)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(a, a.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(a, ngpu):
    edge = weak_edge(a.size, ngpu)
    return {"workload": "%s3d %d^3 = %d DOFs, %d subdomains (%d per GPU), METIS dual, geneo %s tau=%s, %s rtol %g" % (
        a.kind, edge, edge ** 3, a.subs_per_gpu * ngpu, a.subs_per_gpu, a.lvl, a.tau, a.ksp, a.rtol),
        "generator": gen_args(a, edge), "l2": "inputs_larger_than_L2 (factors >> 126 MB)"}


def weak_edge(size, g):
    # the reference's own weak-scaling rule incl. its float truncation (tst/laplacian/laplacian.cpp:104)
    import math
    return size if g == 1 else int(math.floor((float(size) ** 3 * g) ** (1.0 / 3.0) + 1e-9))


# ---------------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------------
def run_b200(a):
    import numpy as np
    import torch
    import geneo4petsc_b200 as g
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        from geneo4petsc_b200 import dist
        return dist.run_bench(a, rank, world, local, METRIC, UNIT, workload_config(a, world), ClockSampler)

    edge = a.size
    t0 = time.time()
    prob = g.Problem().generate(a.kind, gen_args(a, edge))
    t1 = time.time()
    prob.decompose(a.subs_per_gpu, True, 0)
    t2 = time.time()
    n = prob.sizes()["nb_node"]
    opts = ["-geneo_lvl", a.lvl, "-geneo_tau", a.tau, "-geneo_kernel_timing"]
    pc = g.GeneoPC(opts)
    pc.setup(prob)  # cold: host symbolic + upload + numeric
    tm_cold = pc.timers()
    st = pc.stats()
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    b = torch.zeros(n, dtype=torch.float64, device="cuda")
    ones = torch.arange(1, n + 1, dtype=torch.float64, device="cuda")
    pc.mult_device(ones.data_ptr(), b.data_ptr())  # b = A (1..N), src/geneo4PETSc.cpp:820-831
    torch.cuda.synchronize()

    def step():
        pc.refactor()
        return pc.ksp_solve_device(b.data_ptr(), x.data_ptr(), ksp=a.ksp, rtol=a.rtol, atol=1e-50, restart=30)

    for _ in range(a.warmup):
        r = step()
    torch.cuda.synchronize()
    pc.kernel_time()
    c0 = g.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    setup_s = iter_s = 0.0
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.steps):
            ta = time.perf_counter()
            pc.refactor()
            tb = time.perf_counter()
            r = pc.ksp_solve_device(b.data_ptr(), x.data_ptr(), ksp=a.ksp, rtol=a.rtol, atol=1e-50, restart=30)
            tc = time.perf_counter()
            setup_s += tb - ta
            iter_s += tc - tb
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    c1 = g.counters()
    kms, klaunch = pc.kernel_time()
    assert r["reason"] > 0, "KSP did not converge: %s" % r["reason_name"]
    err = float((x - ones).abs().max() / n)
    value = n / (ms * 1e-3)

    # PC-apply and SpMV alone (device pointers), algorithmic GB/s
    rates = {}
    y = torch.empty_like(x)
    for name, fn, nbytes in (("pc_apply", pc.apply_device, st["apply_bytes"]), ("spmv", pc.mult_device, st["spmv_bytes"])):
        for _ in range(3):
            fn(b.data_ptr(), y.data_ptr())
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(10):
            fn(b.data_ptr(), y.data_ptr())
        f1.record()
        torch.cuda.synchronize()
        t = f0.elapsed_time(f1) / 10
        rates[name] = {"ms": t, "GBps": nbytes / t / 1e6}
    pc.kernel_time()

    # end to end through the C ABI with host buffers
    bh = b.cpu().numpy()
    e2e_t, h2d, d2h = [], 0, 0
    tm2 = {"symbolic": 0.0, "upload": 0.0, "numeric": 0.0}
    for _ in range(a.e2e_steps):
        k0 = g.counters()
        torch.cuda.synchronize()
        ta = time.perf_counter()
        pc2 = g.GeneoPC(["-geneo_lvl", a.lvl, "-geneo_tau", a.tau])
        pc2.setup(prob)
        r2 = pc2.ksp_solve(bh, ksp=a.ksp, rtol=a.rtol, atol=1e-50, restart=30)
        tb = time.perf_counter()
        k1 = g.counters()
        e2e_t.append(tb - ta)
        h2d, d2h = k1["h2d"] - k0["h2d"], k1["d2h"] - k0["d2h"]
        tm2 = pc2.timers()
        assert r2["reason"] > 0
        del pc2
    e2e_s = sum(e2e_t) / len(e2e_t) if e2e_t else float("inf")

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    ach = st["trisolve_bytes"] / (kms / max(1, klaunch)) / 1e6 if klaunch else 0.0
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of THIS workload
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = tr.get(workload_config(a, 1)["workload"])
    except (OSError, ValueError):
        pass
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, 1),
        "clocks": clk.summary(),
        "e2e": {"value": n / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "seconds": e2e_s,
                "symbolic_s": tm2["symbolic"], "upload_s": tm2["upload"], "numeric_s": tm2["numeric"]},
        "gpu_launches": c1["launches"] - c0["launches"],
        "roofline": {"kernel": "k_solve_ring<1> (level-1 triangular sweeps of all subdomains, one launch per PC apply)",
                     "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (sustained copy)" if peaks else "fallback 6650",
                     "traffic": traffic, "algorithmic_bytes_per_launch": st["trisolve_bytes"], "launches_timed": klaunch,
                     "avg_launch_ms": kms / max(1, klaunch)},
        "detail": {"n_dof": n, "iterations": r["its"], "reason": r["reason_name"], "rnorm": r["rnorm"], "max_rel_err_vs_1..N": err,
                   "setup_numeric_s": setup_s / a.steps, "iter_s": iter_s / a.steps, "dimE": pc.info()["nE"],
                   "cold_setup": {k: tm_cold[k] for k in ("symbolic", "upload", "numeric", "operator", "setup")},
                   "numeric_phases_s": {k: pc.timers()[k] for k in ("lvl1SetupMinv", "lvl2SetupSyl", "lvl2SetupEig", "lvl2SetupZ", "lvl2SetupE")},
                   "gen_s": t1 - t0, "part_decomp_s": t2 - t1, "factor_bytes": st["factor_bytes"], "factor_flops": st["factor_flops"],
                   "factor_TFLOPs_l1": st["factor_flops"] / max(tm_cold["lvl1SetupMinv"], 1e-9) / 1e12,
                   "pc_apply": rates["pc_apply"], "spmv": rates["spmv"], "hbm_peak_GBps": peak},
    }
    if not a.no_cpu_baseline:
        nn, secs, its, cores = cpu_sample(a, a.cpu_size, a.subs_per_gpu)
        out["cpu_baseline"] = {"value": nn / secs, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "%s %d^3 = %d DOFs, %d subdomains, same options (%d its, %.1f s): scipy SuperLU/ARPACK "
                                         "restatement of the reference (PETSc/MUMPS/SLEPc absent)" % (a.kind, a.cpu_size, nn, a.subs_per_gpu, its, secs)}
    print(json.dumps(out))
    if os.environ.get("GENEO_PROFILE"):
        g.profile_dump(os.environ.get("GENEO_PROFILE_OUT", "gpurun_out/profile_sites.csv"))


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
