#!/usr/bin/env python
"""bench.py -- ADMM throughput of the MultivarTV hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg4|cfg5|small]

A "step" is ONE ADMM pass (x-update by matrix-free PCG + fused z/u update) over the whole synthetic mesh.
metric = mesh-vertex-updates/sec = N_vertices * passes / seconds (admm_iters_per_sec is reported beside it).

Default workload = BASELINE.json north_star target: configs[2], the 3-D 512^3 mesh with n = 2^26 points, fp64, RCPP
mode; `--gpus N` slab-shards THAT mesh over N GPUs (strong scaling).  `value` is timed with the operators (Oty, counts)
already resident in HBM, CUDA events on the plan's stream, per-kernel profiling OFF, max over ranks; the per-kernel
`stages` / `roofline` come from a second, shorter run with the profiling events on.  `e2e` goes through the
reference-shaped call with HOST buffers (points in, theta / fitted out).  Before anything is timed, `parity` runs a
small case through the SAME N-rank plan path and compares it with the CPU oracle (identical Counter, 1e-9): a failure
ends the run with a non-zero exit code.  At N = 1 the line also carries blocks for the other BASELINE configs
(`classical_cg`: the same workload with the Jacobi preconditioner, i.e. the classical CG that SURVEY 8(d)'s byte formula
assumes; `cfg2`, `cfg4`, `cfg5`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]
    "cfg2": dict(m=[4096, 4096], n=1 << 24, desc="2-D 4096x4096 mesh, n=16Mi synthetic points, fp64"),
    # configs[2]: the north_star target
    "cfg3": dict(m=[512, 512, 512], n=1 << 26, desc="3-D 512^3 mesh, n=64Mi synthetic points, fp64"),
    # configs[3]
    "cfg4": dict(m=[96, 96, 96, 96], n=96 ** 4, desc="4-D 96^4 mesh, n=N synthetic points"),
    # configs[4]
    "cfg5": dict(m=[256, 256, 256], n=256 ** 3, desc="3-D 256^3 mesh, n=N synthetic points, 32-lambda path"),
    "small": dict(m=[64, 64, 64], n=1 << 17, desc="3-D 64^3 mesh (smoke-sized)"),
    "tiny": dict(m=[16, 16, 12], n=2000, desc="3-D 16x16x12 mesh (dry runs on the CPU emulator)"),
}
# bounded CPU samples: the same generator / point density / lambda / mode on a smaller mesh of the same family
CPU_SAMPLES = {
    "cfg2": dict(m=[1024, 1024], n=1 << 20),
    "cfg3": dict(m=[128, 128, 128], n=1 << 20),
    "cfg4": dict(m=[32, 32, 32, 32], n=32 ** 4),
    "cfg5": dict(m=[128, 128, 128], n=128 ** 3),
    "small": dict(m=[32, 32, 32], n=1 << 14),
    "tiny": dict(m=[8, 8, 8], n=300),
}
NOMINAL_HBM_GBS = 8000.0   # north_star's "8 TB/s peak"


def synth_points(n, p, seed, z_lo=0.0, z_hi=1.0):
    """Seeded synthetic noisy step function on U(0,1)^p; the last coordinate is drawn in
    [z_lo, z_hi) so a rank can generate exactly the points of its own slab (weak scaling)."""
    rng = np.random.default_rng(seed)
    x = rng.random((n, p))
    if z_lo != 0.0 or z_hi != 1.0:
        x[:, -1] = z_lo + (z_hi - z_lo) * x[:, -1]
    f = np.prod(x > 0.5, axis=1) * 1.0 + 0.5 * np.prod(x < 0.2, axis=1)
    y = f + 0.5 * rng.standard_normal(n)
    return x, y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = getattr(self, "t0", 0.0), getattr(self, "t1", float("inf"))
        inside = [ln for (ts, ln) in self.lines if t0 - 0.02 <= ts <= t1 + 0.12]   # samples taken DURING the timed region
        if not inside:
            inside = [ln for (ts, ln) in self.lines if ts >= t0 - 0.02] or [ln for (_, ln) in self.lines]
        for ln in inside:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_threads():
    """All host threads of the box, whatever OMP_NUM_THREADS says (torch.distributed.run sets it to 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def stage_totals(N, R, esz, passes, inner, deg, kernels):
    """Algorithmic bytes of each kernel class over a run of `passes` ADMM passes with `inner` CG iterations of polynomial
    degree `deg` (0 = Jacobi), DESIGN.md 'Kernels': N = vertices, R = rows of D (local slab)."""
    pw = kernels.get("cg_prec_words", 4)                 # first preconditioner pass: 3 N (diag(c) derived from dinv) or 4 N
    fused = bool(kernels.get("fused_update")) and deg >= 1
    rest = max(0, deg - 1) * 4 * N                       # Horner passes 2..deg: read w, dinv, r ; write w
    t = {
        "zu": passes * (2 * R + 4 * N),                  # read u, theta, theta_prev ; write u, D^T alpha, D^T u
        "cg_init": passes * 8 * N,                       # read theta, c, dinv, Oty, v1, v2 ; write r, theta_old
        "cg_step": inner * (5 if deg else 6) * N,        # read z | (r, dinv), p_old, c ; write p_new, q
    }
    if fused:
        t["cg_update"] = inner * 8 * N                   # read theta, p, r, q, dinv ; write theta, r, w_1
        t["cg_prec"] = passes * (3 * N + rest) + inner * rest   # stand-alone passes before the first iteration + Horner rest
    else:
        t["cg_update"] = inner * (6 if deg else 7) * N   # read theta, p, r, q (, dinv) ; write theta, r
        t["cg_prec"] = inner * (pw * N + rest) if deg else 0
    # launch groups that did work (launches of a batch that come after convergence return at once)
    perf = {"zu": passes, "cg_init": passes, "cg_step": inner, "cg_update": inner,
            "cg_prec": (passes + (inner if deg >= 2 else 0)) if fused else (inner if deg else 0)}
    return {k: esz * v for k, v in t.items()}, perf


def cpu_baseline(workload, lam, cg_rtol, threads, passes=12, warm=1):
    """The oracle port (oracle/c/mvtv_oracle.c, matrix-free, Jacobi-PCG x-update, OpenMP) on a bounded sample of the
    workload's family, with an explicit thread count."""
    from oracle import c_oracle as co
    smp = CPU_SAMPLES[workload]
    m, n = smp["m"], smp["n"]
    x, y = synth_points(n, len(m), 117)
    axes = [np.linspace(0.0, 1.0, d) for d in m]
    kw = dict(mode=co.MODE_RCPP, solver=co.SOLVER_PCG, cg_rtol=cg_rtol, nthreads=threads)
    if warm:
        co.mbs_one(x, y, m, axes, lam, max_passes=warm, **kw)
    r = co.mbs_one(x, y, m, axes, lam, max_passes=passes, **kw)
    N = int(np.prod(m))
    return dict(value=N * r["passes"] / r["seconds"], unit="vertex-updates/s", cores=threads, kind="port",
                sample="oracle/c (matrix-free C port, OpenMP %d threads, Jacobi-PCG rtol %.0e) on a %s mesh, n=%d, %d ADMM "
                       "passes after %d warm-up, same generator/lambda/mode as the workload" % (
                           threads, cg_rtol, "x".join(map(str, m)), n, r["passes"], warm),
                seconds=r["seconds"], passes=r["passes"], inner_cg_iters=r["inner_iters"], mesh=m, n_points=n)


def reference_compiled_config1():
    """The reference's OWN compiled cpp-code solver (oracle/_ref, built from /root/reference against the Armadillo stand-in)
    on BASELINE configs[0] (2-D, n = 1000 points, 32 x 32 mesh, lambda = 1.5) -- informational: upstream's O(n N) nearest
    search and per-pass factorisation make it unusable at the benchmark's sizes."""
    try:
        from oracle import ref_oracle as ro
        if not (os.path.exists(ro.LIB) and os.path.exists(ro.LIB_RCPP)):
            return None
        from tests.helpers import synth
        x, y = synth(117, 1000, 2)
        t0 = time.perf_counter()
        r = ro.mbs_one(x, y, [32, 32], 1.5)
        dt = time.perf_counter() - t0
        passes = r["counter"] - 1
        return {"what": "cpp-code mbs_one (set-up + %d ADMM passes) on configs[0], single thread" % passes, "seconds": dt,
                "counter": r["counter"], "vertex_updates_per_s": 1024 * passes / dt}
    except Exception as e:   # informational only
        return {"error": str(e)[:200]}


def run_reference(args):
    """--impl reference: the reference's CPU path.  Upstream's own code (oracle/_ref) cannot run the benchmark's sizes
    (O(n N) nearest search, a sparse LU per pass), so the timed value is the oracle port with ALL host threads of the box
    (explicit count: torchrun's OMP_NUM_THREADS=1 is ignored) on the bounded sample of the workload's family -- the line's
    `config` names the mesh that was really timed, `sample_of` the workload it stands for.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    threads = host_threads()
    cb = cpu_baseline(args.workload, args.lam, args.cg_rtol, threads, passes=args.steps, warm=args.warmup)
    line = {
        "impl": "reference", "metric": "mesh_vertex_updates_per_sec", "value": cb["value"], "unit": "vertex-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * cb["seconds"] / max(1, cb["passes"]), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "bounded CPU sample of: " + wl["desc"], "mesh": cb["mesh"], "n_points": cb["n_points"],
                   "mode": "rcpp", "lambda": args.lam, "cg_rtol": args.cg_rtol, "precond": "jacobi",
                   "sample_of": {"workload": wl["desc"], "mesh": wl["m"], "n_points": wl["n"]}, "same_config": False},
        "admm_iters_per_sec": cb["passes"] / cb["seconds"],
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "vertex-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "sample_vertices": int(np.prod(cb["mesh"])), "reference_compiled_config1": reference_compiled_config1(),
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
def make_plan(mv, m, dtype, dist, rank, world, local_rank):
    uid = None
    if world > 1:
        box = [mv.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    return mv.Plan(m, dtype=dtype, device=local_rank, rank=rank, world=world, nccl_unique_id=uid)


def slab_points(x, y, axes, rank, world):
    if world == 1:
        return x, y
    from multivartv_b200 import partition
    keep = partition.owner_of(x[:, -1], axes[-1], world) == rank
    return x[keep], y[keep]


def parity_block(mv, dist, rank, world, local_rank, precond, tag):
    """A small solve through the SAME N-rank plan path (slab partition, peer / NCCL collectives, the preconditioner the
    timed run uses) against the single-process CPU oracle: identical Counter, max|dtheta| <= 1e-9."""
    from oracle import c_oracle as co
    m, n, passes, lam = [40, 40, 40], 64000, 15, 1.0
    x, y = synth_points(n, 3, 41)
    axes = [np.linspace(0.0, 1.0, d) for d in m]
    xs, ys = slab_points(x, y, axes, rank, world)
    plan = make_plan(mv, m, mv.F64, dist, rank, world, local_rank)
    plan.set_points(xs, ys, axes)
    out = plan.solve(lam, mode="rcpp", max_passes=passes, want_fitted=False, cg_rtol=1e-13, precond=precond)
    desc = plan.describe()
    plan.close()
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (rank, out["theta"], out["counter"], out["inner_iters"]))
    else:
        gathered = [(0, out["theta"], out["counter"], out["inner_iters"])]
    res = None
    if rank == 0:
        gathered.sort(key=lambda t: t[0])
        theta = np.concatenate([g[1] for g in gathered])
        ref = co.mbs_one(x, y, m, axes, lam, mode=co.MODE_RCPP, max_passes=passes, solver=co.SOLVER_PCG, cg_rtol=1e-13,
                         nthreads=host_threads())
        err = float(np.abs(theta - ref["theta"]).max())
        same = all(g[2] == ref["counter"] for g in gathered)
        res = {"case": "3-D 40^3 mesh, n=64000, %d RCPP passes, precond=%s, world=%d" % (passes, tag, world),
               "counter_equal": bool(same), "counter": int(gathered[0][2]), "oracle_counter": int(ref["counter"]),
               "max_abs": err, "tol": 1e-9, "ok": bool(same and err <= 1e-9), "inner_cg_iters": int(gathered[0][3]),
               "kernels": {k: desc[k] for k in ("cg_step", "collectives", "fold_commit", "last_degree")}}
    if world > 1:
        box = [res]
        dist.broadcast_object_list(box, src=0)
        res = box[0]
    return res


def timed_passes(mv, plan, args, precond, steps, warm_passes, dist, profile=False, sampler=None):
    """ONE solve call from a cold start: W untimed warm-up passes, then exactly `steps` passes timed on the device (CUDA
    events on the plan's stream around passes W+1 .. W+K, `timing_skip_passes`), so that no timed pass is the first pass of
    a call.  Returns the result (timed_* fields = the timed interval), max-over-ranks device and wall seconds, and the
    per-class profile of the WHOLE call when profiling is on."""
    kw = dict(mode=args.mode, cg_rtol=args.cg_rtol, want_theta=False, want_fitted=False, raise_on_nonconvergence=False,
              precond=precond)
    W = max(1, warm_passes)
    if dist:
        import torch
        torch.cuda.synchronize()
        dist.barrier()
    plan.profile(profile)
    if sampler:
        sampler.begin()
    t0 = time.perf_counter()
    r = plan.solve(args.lam, max_passes=W + steps, timing_skip_passes=W, **kw)
    wall = time.perf_counter() - t0
    if sampler:
        sampler.end()
    prof = plan.get_profile() if profile else None
    plan.profile(False)
    dev_s = r["device_seconds"]
    if dist:
        import torch
        torch.cuda.synchronize()
        t = torch.tensor([dev_s, wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, wall = float(t[0]), float(t[1])
    if r["timed_passes"] != steps:
        raise SystemExit("bench: solver stopped after %d of %d passes (converged early?); pick another lambda" % (r["passes"], W + steps))
    return r, dev_s, wall, prof


def pass_fractions(Nl, Rl, esz, J, ms_per_pass, peak, moved_bytes_per_pass):
    """B_pass = T [(2R + 3N) + 12 N J] with the J actually run, against the measured and the nominal peak."""
    b = esz * ((2 * Rl + 3 * Nl) + 12 * Nl * J)
    gbs = b / (ms_per_pass * 1e-3) / 1e9
    kg = moved_bytes_per_pass / (ms_per_pass * 1e-3) / 1e9
    return {"formula": "T*((2R+3N) + 12*N*J), J = CG iterations per pass actually run (SURVEY 8(d))",
            "alg_bytes": b, "gbs": gbs, "frac_of_measured_peak": gbs / peak, "frac_of_8TBs": gbs / NOMINAL_HBM_GBS,
            "kernel_bytes": moved_bytes_per_pass, "kernel_gbs": kg, "kernel_frac_of_measured_peak": kg / peak,
            "kernel_frac_of_8TBs": kg / NOMINAL_HBM_GBS}


def run_workload(mv, args, name, dist, rank, world, local_rank, precond, tag, steps, warmup, dtype="f64",
                 with_stages=True, with_e2e=True, sampler=None):
    """One workload through the plan API: returns the fields of a bench line (rank 0) or None."""
    wl = WORKLOADS[name]
    m = list(wl["m"])
    p = len(m)
    n = wl["n"]
    esz = 8 if dtype == "f64" else 4
    if world > 1 and args.scaling == "weak":
        m[-1] *= world   # every rank keeps a full workload-sized slab
    N = int(np.prod(m))
    axes = [np.linspace(0.0, 1.0, d) for d in m]
    plan = make_plan(mv, m, mv.F64 if dtype == "f64" else mv.F32, dist, rank, world, local_rank)

    t_gen = time.time()
    if world == 1:
        x, y = synth_points(n, p, 117)
    elif args.scaling == "weak":
        from multivartv_b200 import partition
        lo, hi = partition.slab_interval(axes[-1], plan.z0, plan.nz)
        x, y = synth_points(n, p, 117 + rank, lo, hi)
    else:
        xa, ya = synth_points(n, p, 117)
        x, y = slab_points(xa, ya, axes, rank, world)
        del xa, ya
    t_gen = time.time() - t_gen
    n_local = x.shape[0]
    # host side of the e2e call: inputs and outputs live in page-locked memory (allocated outside any timed region)
    xp, yp = mv.pinned_empty(x.shape), mv.pinned_empty(y.shape)
    xp[...] = x
    yp[...] = y
    x, y = xp, yp
    plan.set_points(x, y, axes)

    # ---- value: exactly K passes, inputs resident in HBM, profiling off ------------------------------
    r, dev_s, wall, _ = timed_passes(mv, plan, args, precond, steps, warmup, dist, profile=False, sampler=sampler)
    passes, inner, launches = r["timed_passes"], r["timed_inner_iters"], r["timed_kernel_launches"]
    kernels = plan.describe()
    deg = kernels["last_degree"]

    # ---- stages: a second, shorter run with the per-kernel CUDA events on ----------------------------
    stages, prof_run = {}, None
    if with_stages:
        ps = max(2, min(steps, 5))
        rp, dev_p, _, prof = timed_passes(mv, plan, args, precond, ps, warmup, dist, profile=True)
        prof_run = {"passes": rp["passes"], "ms_per_step_with_profiling_events": 1e3 * dev_p / ps, "inner_cg_iters": rp["inner_iters"],
                    "note": "per-class totals cover the whole call (warm-up passes included)"}
        Nl, Rl = plan.n_local, plan.R * plan.n_local / max(1, plan.N)
        tot, perf = stage_totals(Nl, Rl, esz, rp["passes"], rp["inner_iters"], deg, kernels)
        for k, (ms, cnt) in prof.items():
            if k in tot and cnt > 0 and tot[k] > 0 and ms > 0 and perf[k] > 0:
                stages[k] = {"total_ms": ms, "launch_groups": cnt, "performed": perf[k], "avg_ms": ms / perf[k],
                             "alg_bytes_total": tot[k], "alg_bytes_per_launch_group": tot[k] / perf[k],
                             "gbs": tot[k] / (ms * 1e-3) / 1e9}

    # ---- e2e: host buffers in, theta + fitted out, through the public call ------------------------
    e2e = None
    if with_e2e:
        theta_host, fitted_host = mv.pinned_empty(plan.n_local), mv.pinned_empty(n_local)
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        plan.set_points(x, y, axes)
        t_sp = time.perf_counter() - t0
        re = plan.solve(args.lam, max_passes=steps, mode=args.mode, cg_rtol=args.cg_rtol, want_theta=True,
                        want_fitted=True, raise_on_nonconvergence=False, precond=precond, theta_out=theta_host,
                        fitted_out=fitted_host)
        t_e2e = time.perf_counter() - t0
        if dist:
            import torch
            t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_e2e = float(t[0])
        h2d = 8 * (n_local * (p + 1) + sum(m))
        d2h = 8 * (plan.n_local + n_local)
        e2e = {"value": N * re["passes"] / t_e2e, "unit": "vertex-updates/s",
               "h2d_bytes_per_step": h2d / max(1, re["passes"]), "d2h_bytes_per_step": d2h / max(1, re["passes"]),
               "call": "Plan.set_points(pinned host x,y) + Plan.solve(%d passes, cold start) -> pinned host theta, fitted" % re["passes"],
               "seconds": t_e2e, "passes": re["passes"], "set_points_seconds": t_sp,
               "solve_device_seconds": re["device_seconds"], "inner_cg_iters": re["inner_iters"]}

    Nl, Rl = plan.n_local, plan.R * plan.n_local / max(1, plan.N)
    plan.close()   # every rank drops its plan (and NCCL communicator) at the same point: ncclCommDestroy is collective
    if rank != 0:
        return None
    peak, peak_src = measured_peak_gbs()
    J = inner / max(1, passes)
    moved = sum(stage_totals(Nl, Rl, esz, 1, J, deg, kernels)[0].values())
    for s in stages.values():
        s["frac"] = s["gbs"] / peak
    dom = max(stages, key=lambda k: stages[k]["total_ms"]) if stages else None
    roof = None
    if dom:
        tot_ms = sum(s["total_ms"] for s in stages.values())
        fam = kernels["cg_step"]
        strip = fam != "k_cg_step"
        fusedk = bool(kernels.get("fused_update")) and deg >= 1
        kname = {"zu": kernels["zu"], "cg_init": (fam + "<STEP_INIT>") if fam == "k_cg_step3d" else ("k_cg_init2d" if strip else "k_cg_init"),
                 "cg_step": fam + ("<STEP_Z>" if deg else "<STEP_JACOBI>"),
                 "cg_update": ((fam + "<STEP_UPDPREC>") if fam == "k_cg_step3d" else "k_cg_updprec2d") if fusedk else "k_cg_update",
                 "cg_prec": (((fam + "<STEP_HORNER>") if fam == "k_cg_step3d" else "k_cg_horner2d") + " x %d per CG iteration" % (deg - 1)) if (fusedk and deg >= 2)
                            else fam + "<STEP_PREC>"}[dom]
        # launches per launch group of the dominant class: the Horner passes 2..d of an iteration are timed as one group
        lpg = (deg - 1) if (dom == "cg_prec" and fusedk and deg >= 2) else 1
        traffic = ncu_traffic(name, dom, kernels, world, dtype)
        roof = {"bound": "hbm", "kernel": kname, "kernel_class": dom, "achieved": stages[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": stages[dom]["frac"], "traffic": (traffic / lpg) if traffic else None,
                "alg_bytes_per_launch": stages[dom]["alg_bytes_per_launch_group"] / lpg, "avg_launch_ms": stages[dom]["avg_ms"] / lpg,
                "launches_per_cg_iteration": lpg,
                "peak_source": peak_src, "share_of_step": stages[dom]["total_ms"] / max(tot_ms, 1e-9),
                "pass": pass_fractions(Nl, Rl, esz, J, 1e3 * dev_s / passes, peak, moved)}
    return {
        "value": N * passes / dev_s, "ms_per_step": 1e3 * dev_s / passes,
        "config": {"workload": wl["desc"] + (" per GPU (weak: last axis x%d)" % world if world > 1 and args.scaling == "weak" else ""),
                   "mesh": m, "n_points": n * (world if args.scaling == "weak" else 1), "mode": args.mode, "lambda": args.lam,
                   "cg_rtol": args.cg_rtol, "precond": tag, "poly_degree": deg, "parallelism": "slab%d" % world,
                   "kernels": {k: kernels[k] for k in ("zu", "cg_step", "fused_update", "collectives", "fold_commit")},
                   "l2": "working set per pass (u: %.2f GB) exceeds the 126 MB L2" % (2 * Rl * esz / 1e9)},
        "admm_iters_per_sec": passes / dev_s, "inner_cg_iters_per_pass": J,
        "wall_seconds": wall, "device_seconds": dev_s, "gen_seconds": t_gen, "gpu_launches": launches,
        "timing": "one solve call from a cold start: %d untimed warm-up passes, then %d passes between two CUDA events on the plan's stream" % (max(1, warmup), steps),
        "roofline": roof, "stages": stages, "stages_run": prof_run, "e2e": e2e,
    }


def ncu_traffic(workload, dom, kernels, world, dtype):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this workload."""
    if world != 1 or dtype != "f64":
        return None
    for tf in ("r2_ncu_traffic.json",):
        try:
            ent = json.load(open(os.path.join(ROOT, "profiles", tf))).get(workload, {}).get(dom)
        except Exception:
            ent = None
        if ent:
            return ent["dram_bytes_per_launch"]
    return None


def small_block(d):
    """The fields of a secondary block (other configs / preconditioners riding on the main line)."""
    keys = ("value", "ms_per_step", "admm_iters_per_sec", "inner_cg_iters_per_pass", "gpu_launches")
    out = {k: d[k] for k in keys}
    out["config"] = {k: d["config"][k] for k in ("workload", "mesh", "n_points", "precond", "poly_degree")}
    if d.get("roofline"):
        out["pass"] = d["roofline"]["pass"]
        out["dominant_kernel"] = {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "share_of_step")}
    out["stages"] = {k: {"avg_ms": v["avg_ms"], "gbs": v["gbs"], "frac": v["frac"]} for k, v in d.get("stages", {}).items()}
    return out


def cfg4_precision_check(mv, args):
    """BASELINE configs[3] 'fp64 vs fp32 tolerance check': fp32 and fp64 plans on the full 96^4 mesh after a bounded number
    of passes, plus both against the CPU ORACLE on a 4-D mesh the oracle finishes in seconds (fp32 <= 1e-4, fp64 <= 1e-9)."""
    from oracle import c_oracle as co
    out = {}
    m, n, passes = [20, 20, 20, 20], 160000, 12
    x, y = synth_points(n, 4, 117)
    axes = [np.linspace(0.0, 1.0, d) for d in m]
    ref = co.mbs_one(x, y, m, axes, args.lam, mode=co.MODE_RCPP, max_passes=passes, solver=co.SOLVER_PCG, cg_rtol=1e-13,
                     nthreads=host_threads())
    for nm, dt in (("f64", mv.F64), ("f32", mv.F32)):
        with mv.Plan(m, dtype=dt) as pl:
            pl.set_points(x, y, axes)
            o = pl.solve(args.lam, mode="rcpp", max_passes=passes, want_fitted=False, raise_on_nonconvergence=False)
        out["oracle_20^4_" + nm] = {"max_abs": float(np.abs(o["theta"] - ref["theta"]).max()), "passes": o["passes"],
                                    "counter_equal": bool(o["counter"] == ref["counter"]), "tol": 1e-9 if nm == "f64" else 1e-4}
    return out


def cfg5_path(mv, args):
    """BASELINE configs[4]: 32 log-spaced lambdas on the 256^3 mesh (n = N), warm-started, without leaving the device.
    CPP mode runs every lambda to the reference's own stopping rule (upstream's `int rho` ends each after a few passes);
    RCPP mode is capped at 10 passes per lambda to keep the default run short."""
    wl = WORKLOADS["cfg5"]
    m, n = wl["m"], wl["n"]
    N = int(np.prod(m))
    x, y = synth_points(n, 3, 117)
    axes = [np.linspace(0.0, 1.0, d) for d in m]
    lams = np.exp(np.linspace(np.log(10.0), np.log(10.0 * 1e-4), 32))   # flipud(exp(linspace)) convention: descending
    out = {"config": {"workload": wl["desc"], "mesh": m, "n_points": n, "lambdas": [float(lams[0]), float(lams[-1])], "n_lambda": 32}}
    with mv.Plan(m) as pl:
        pl.set_points(x, y, axes)
        for mode, cap in (("cpp", 0), ("rcpp", 10)):
            t0 = time.perf_counter()
            r = pl.solve_path(lams, y, mode=mode, max_counter=(cap + 1 if cap else 0), want_best=False)
            dt = time.perf_counter() - t0
            out[mode] = {"path_seconds": dt, "device_seconds": r["device_seconds"], "passes": int(r["passes"]),
                         "inner_cg_iters": int(r["inner_iters"]), "vertex_updates_per_s": N * r["passes"] / max(r["device_seconds"], 1e-9),
                         "counters_first_last": [int(r["counters"][0]), int(r["counters"][-1])],
                         "passes_cap_per_lambda": cap or None}
    return out


def run_ours(args):
    import multivartv_b200 as mv
    from multivartv_b200 import build as mvbuild

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=300))
    if rank == 0:
        mvbuild.build()
    if dist:
        dist.barrier()
    PRE = {"auto": mv.PRECOND_AUTO, "jacobi": mv.PRECOND_JACOBI, "cheb1": mv.PRECOND_CHEB1, "cheb2": mv.PRECOND_CHEB2,
           "cheb3": mv.PRECOND_CHEB3, "cheb4": mv.PRECOND_CHEB4}
    precond = PRE[args.precond]

    # ---- parity first: the same N-rank path against the CPU oracle --------------------------------
    parity = None
    if not args.no_parity:
        parity = parity_block(mv, dist, rank, world, local_rank, precond, args.precond)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": "mesh_vertex_updates_per_sec", "value": None, "parity": parity,
                                  "error": "parity check failed: nothing was timed"}), flush=True)
            if dist:
                dist.barrier()
                dist.destroy_process_group()
            raise SystemExit(3)

    sampler = ClockSampler(local_rank)
    sampler.start()            # nvidia-smi needs a moment to come up: start it before the warm-up
    main = run_workload(mv, args, args.workload, dist, rank, world, local_rank, precond, args.precond, args.steps, args.warmup,
                        dtype=args.dtype, with_e2e=not args.no_e2e, sampler=sampler)
    clocks = sampler.stop()
    if rank != 0:
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return
    line = {
        "metric": "mesh_vertex_updates_per_sec", "value": main["value"], "unit": "vertex-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"],
        "higher_is_better": True, "scaling": args.scaling if world > 1 else "strong", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic", "config": main["config"],
        "admm_iters_per_sec": main["admm_iters_per_sec"], "inner_cg_iters_per_pass": main["inner_cg_iters_per_pass"],
        "wall_seconds": main["wall_seconds"], "device_seconds": main["device_seconds"], "gen_seconds": main["gen_seconds"],
        "timing": main["timing"],
        "clocks": clocks, "gpu_launches": main["gpu_launches"], "parity": parity, "roofline": main["roofline"],
        "stages": main["stages"], "stages_run": main["stages_run"], "e2e": main["e2e"],
    }
    # ---- secondary blocks (one GPU): classical CG on the same workload, the other BASELINE configs ----
    if world == 1 and not args.no_blocks:
        half = max(3, args.steps // 2)
        try:
            if args.precond != "jacobi":
                line["classical_cg"] = small_block(run_workload(mv, args, args.workload, None, 0, 1, local_rank, mv.PRECOND_JACOBI,
                                                                "jacobi", half, args.warmup, with_e2e=False))
                # SURVEY 8(d)'s byte formula describes classical (Jacobi) CG: its fraction for that configuration sits beside the
                # default's, whose polynomial preconditioner lowers J -- and with it both the time per pass and the formula's bytes
                if line.get("roofline") and line["classical_cg"].get("pass"):
                    line["roofline"]["pass_classical_cg"] = dict(line["classical_cg"]["pass"], ms_per_step=line["classical_cg"]["ms_per_step"],
                                                                 inner_cg_iters_per_pass=line["classical_cg"]["inner_cg_iters_per_pass"])
            for other in ("cfg2", "cfg4"):
                if other != args.workload and args.workload == "cfg3":
                    line[other] = small_block(run_workload(mv, args, other, None, 0, 1, local_rank, precond, args.precond,
                                                           half, args.warmup, with_e2e=False))
            if args.workload == "cfg3":
                line["cfg4"]["precision_check"] = cfg4_precision_check(mv, args)
                line["cfg5"] = cfg5_path(mv, args)
        except Exception as e:   # a secondary block must not cost the headline
            line["blocks_error"] = repr(e)[:300]
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_baseline(args.workload, args.lam, args.cg_rtol, host_threads(), passes=12, warm=1)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--mode", default="rcpp", choices=["rcpp", "cpp", "py"])
    ap.add_argument("--lam", type=float, default=1.0)
    ap.add_argument("--cg-rtol", dest="cg_rtol", type=float, default=1e-13)
    ap.add_argument("--precond", default="auto", choices=["auto", "jacobi", "cheb1", "cheb2", "cheb3", "cheb4"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-blocks", action="store_true", help="skip classical_cg / cfg2 / cfg4 / cfg5 blocks")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print("bench: warning: fewer than 3 warm-up steps", file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
