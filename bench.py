#!/usr/bin/env python
"""bench.py -- ADMM throughput of the MultivarTV hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4|small]

A "step" is ONE ADMM pass (x-update by matrix-free PCG + fused z/u update) over the whole synthetic mesh.
metric = mesh-vertex-updates/sec = N_vertices * passes / seconds (admm_iters_per_sec is reported beside it).
`value` is timed with the operators (Oty, counts) already resident in HBM (CUDA events on the plan's stream,
max over ranks); `e2e` goes through the reference-shaped call with HOST buffers (points in, theta/fitted out).
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
    # BASELINE.json configs[1]: the configuration the metric is quoted on at N=1
    "cfg2": dict(m=[4096, 4096], n=1 << 24, desc="2-D 4096x4096 mesh, n=16Mi synthetic points, fp64"),
    # configs[2]
    "cfg3": dict(m=[512, 512, 512], n=1 << 26, desc="3-D 512^3 mesh, n=64Mi synthetic points"),
    # configs[3]
    "cfg4": dict(m=[96, 96, 96, 96], n=96 ** 4, desc="4-D 96^4 mesh, n=N synthetic points"),
    "cfg5": dict(m=[256, 256, 256], n=256 ** 3, desc="3-D 256^3 mesh, n=N synthetic points"),
    "small": dict(m=[512, 512], n=1 << 18, desc="2-D 512x512 mesh (smoke-sized)"),
}
CPU_SAMPLE = dict(m=[1024, 1024], n=1 << 20)   # same point density / function / lambda as cfg2, 1/16 of the mesh


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


def stage_bytes(N, R, esz, prec_words=4):
    """Algorithmic bytes per launch of each kernel class (DESIGN.md 'Kernels'): N = vertices, R = rows of D.
    prec_words: 4 for k_cg_step<STEP_PREC> (reads r, dinv, c), 3 for k_cg_step2d (derives c from dinv)."""
    return {
        "zu": esz * (2 * R + 4 * N),      # read u, theta, theta_prev ; write u, D^T alpha, D^T u
        "cg_init": esz * (8 * N),         # read theta, c, dinv, Oty, v1, v2 ; write r, theta_old
        "cg_step": esz * (6 * N),         # read r, dinv, p_old, c ; write p_new, q
        "cg_update": esz * (7 * N),       # read theta, p, r, q, dinv ; write theta, r
        # MVTV_PRECOND_CHEB1 variants
        "cg_prec": esz * (prec_words * N),  # read r, dinv (, c) ; write z
        "cg_step_z": esz * (5 * N),       # read z, p_old, c ; write p_new, q
        "cg_update_p": esz * (6 * N),     # read theta, p, r, q ; write theta, r
    }


def cpu_baseline(args, threads=0, passes=2, warm=0):
    """The oracle port (oracle/c/mvtv_oracle.c, matrix-free, Jacobi-PCG x-update, OpenMP) on a bounded sample."""
    from oracle import c_oracle as co
    m, n = CPU_SAMPLE["m"], CPU_SAMPLE["n"]
    x, y = synth_points(n, len(m), 117)
    axes = [np.linspace(0.0, 1.0, d) for d in m]
    kw = dict(mode=co.MODE_RCPP, solver=co.SOLVER_PCG, cg_rtol=args.cg_rtol, nthreads=threads)
    if warm:
        co.mbs_one(x, y, m, axes, args.lam, max_passes=warm, **kw)
    r = co.mbs_one(x, y, m, axes, args.lam, max_passes=passes, **kw)
    N = int(np.prod(m))
    cores = threads if threads > 0 else (os.cpu_count() or 1)
    return dict(value=N * r["passes"] / r["seconds"], unit="vertex-updates/s", cores=cores, kind="port",
                sample="oracle/c (matrix-free C port, OpenMP, Jacobi-PCG rtol %.0e) on a %s mesh, n=%d, %d ADMM "
                       "passes, same generator/lambda/mode as the workload" % (args.cg_rtol, "x".join(map(str, m)), n,
                                                                                r["passes"]),
                seconds=r["seconds"], passes=r["passes"], inner_cg_iters=r["inner_iters"])


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
    (O(n N) nearest search), so the timed value is the oracle port with all host threads on the bounded sample (rank 0
    only); the compiled reference is timed on BASELINE configs[0] beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    cb = cpu_baseline(args, threads=0, passes=args.steps, warm=args.warmup)
    N = int(np.prod(CPU_SAMPLE["m"]))
    line = {
        "impl": "reference", "metric": "mesh_vertex_updates_per_sec", "value": cb["value"], "unit": "vertex-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * cb["seconds"] / max(1, cb["passes"]), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "mesh": wl["m"], "n_points": wl["n"], "mode": "rcpp", "lambda": args.lam},
        "admm_iters_per_sec": cb["passes"] / cb["seconds"],
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "vertex-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "sample_vertices": N, "reference_compiled_config1": reference_compiled_config1(),
    }
    print(json.dumps(line), flush=True)


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
                                timeout=datetime.timedelta(seconds=180))
    if rank == 0:
        mvbuild.build()
    if dist:
        dist.barrier()

    wl = WORKLOADS[args.workload]
    m = list(wl["m"])
    p = len(m)
    n = wl["n"]
    esz = 8 if args.dtype == "f64" else 4
    dtype = mv.F64 if args.dtype == "f64" else mv.F32
    if world > 1 and args.scaling == "weak":
        m[-1] *= world   # every rank keeps a full workload-sized slab
    N = int(np.prod(m))
    axes = [np.linspace(0.0, 1.0, d) for d in m]

    uid = None
    if world > 1:
        box = [mv.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    plan = mv.Plan(m, dtype=dtype, device=local_rank, rank=rank, world=world, nccl_unique_id=uid)

    # ---- synthetic inputs: each rank generates the points of its own slab ------------------------
    t_gen = time.time()
    if world == 1:
        x, y = synth_points(n, p, 117)
    else:
        from multivartv_b200 import partition
        lo, hi = partition.slab_interval(axes[-1], plan.z0, plan.nz)
        if args.scaling == "weak":
            x, y = synth_points(n, p, 117 + rank, lo, hi)
        else:
            xa, ya = synth_points(n, p, 117)
            keep = partition.owner_of(xa[:, -1], axes[-1], world) == rank
            x, y = xa[keep], ya[keep]
            del xa, ya
    t_gen = time.time() - t_gen
    n_local = x.shape[0]
    # host side of the e2e call: inputs and outputs live in page-locked memory (allocated outside any timed region)
    xp, yp = mv.pinned_empty(x.shape), mv.pinned_empty(y.shape)
    xp[...] = x
    yp[...] = y
    x, y = xp, yp
    theta_host, fitted_host = mv.pinned_empty(plan.n_local), mv.pinned_empty(n_local)

    # ---- warm-up: operators + W passes ------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()            # nvidia-smi needs a moment to come up: start it before the warm-up
    plan.set_points(x, y, axes)
    mode = args.mode
    precond = {"cheb1": mv.PRECOND_CHEB1, "jacobi": mv.PRECOND_JACOBI}[args.precond]
    kw = dict(mode=mode, cg_rtol=args.cg_rtol, want_theta=False, want_fitted=False, raise_on_nonconvergence=False,
              precond=precond)
    rw = plan.solve(args.lam, max_passes=max(1, args.warmup), **kw)
    warm = mv.WARM_THETA_FROM_PLAN | mv.WARM_U_FROM_PLAN

    # ---- timed: exactly K passes, inputs resident in HBM ------------------------------------------
    if dist:
        dist.barrier()
    plan.profile(True)
    sampler.begin()
    t0 = time.perf_counter()
    r = plan.solve(args.lam, max_passes=args.steps, flags=warm, rho_init=rw["rho"], rho_matrix0=rw["rho"], **kw)
    wall = time.perf_counter() - t0
    sampler.end()
    prof = plan.get_profile()
    plan.profile(False)
    clocks = sampler.stop()
    dev_s = r["device_seconds"]
    passes = r["passes"]
    inner = r["inner_iters"]
    launches = r["kernel_launches"]
    if dist:
        import torch
        t = torch.tensor([dev_s, wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, wall = float(t[0]), float(t[1])
    if passes != args.steps:
        raise SystemExit("bench: solver stopped after %d of %d passes (converged early?); pick another lambda"
                         % (passes, args.steps))

    # ---- e2e: host buffers in, theta + fitted out, through the public call ------------------------
    e2e = None
    if not args.no_e2e:
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        plan.set_points(x, y, axes)
        t_sp = time.perf_counter() - t0
        re = plan.solve(args.lam, max_passes=args.steps, mode=mode, cg_rtol=args.cg_rtol, want_theta=True,
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

    # every rank drops its plan (and NCCL communicator) at the same point: ncclCommDestroy is collective
    Nl_, R_, Nfull_ = plan.n_local, plan.R, plan.N
    kernels = plan.describe()
    plan.close()
    if rank != 0:
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    Nl, Rl = Nl_, R_ * Nl_ / max(1, Nfull_)   # rows scale with the slab
    sb = stage_bytes(Nl, Rl, esz, kernels["cg_prec_words"])
    if args.precond == "cheb1":
        sb["cg_step"], sb["cg_update"] = sb["cg_step_z"], sb["cg_update_p"]
    performed = {"zu": passes, "cg_init": passes, "cg_step": inner, "cg_update": inner, "cg_prec": inner}
    stages = {}
    for k, (ms, cnt) in prof.items():
        if k not in sb or cnt == 0 or performed[k] == 0:
            continue
        avg_ms = ms / performed[k]
        stages[k] = {"total_ms": ms, "launches": cnt, "performed": performed[k], "avg_ms": avg_ms,
                     "alg_bytes": sb[k], "gbs": sb[k] / (avg_ms * 1e-3) / 1e9, "frac": sb[k] / (avg_ms * 1e-3) / 1e9 / peak}
    dom = max(stages, key=lambda k: stages[k]["total_ms"]) if stages else None
    # DRAM traffic of the same kernel from the committed ncu --set full capture of this workload (per launch)
    traffic = None
    for tf in ("r1_final2_ncu_traffic.json", "r1_final_ncu_traffic.json"):   # newest capture first
        try:
            ent = json.load(open(os.path.join(ROOT, "profiles", tf))).get(args.workload, {}).get(dom)
        except Exception:
            ent = None
        # only a capture of the kernel that actually ran counts (cg_step / cg_prec exist in two implementations)
        runs = kernels.get(dom, "k_" + str(dom)) + "<"
        if ent and world == 1 and args.dtype == "f64" and ent.get("kernel", "").startswith(runs):
            traffic = ent["dram_bytes_per_launch"]
            break
    roof = None
    if dom:
        roof = {"bound": "hbm", "kernel": "k_" + dom, "achieved": stages[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": stages[dom]["frac"], "traffic": traffic, "alg_bytes_per_launch": stages[dom]["alg_bytes"],
                "peak_source": peak_src,
                "share_of_step": stages[dom]["total_ms"] / (dev_s * 1e3)}
    # whole-pass algorithmic bytes: (2R+3N) + 12 N J
    J = inner / max(1, passes)
    b_iter = esz * ((2 * Rl + 3 * Nl) + 12 * Nl * J)      # SURVEY 8(d) accounting (classical CG: 12 N per inner iteration)
    b_moved = sb["zu"] + sb["cg_init"] + J * (sb["cg_step"] + sb["cg_update"] + (sb["cg_prec"] if args.precond == "cheb1" else 0))
    line = {
        "metric": "mesh_vertex_updates_per_sec", "value": N * passes / dev_s, "unit": "vertex-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / passes,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic",
        "config": {"workload": wl["desc"] + (" per GPU (weak: last axis x%d)" % world if world > 1 and args.scaling == "weak" else ""),
                   "mesh": m, "n_points": n * (world if args.scaling == "weak" else 1), "mode": mode, "lambda": args.lam,
                   "cg_rtol": args.cg_rtol, "precond": args.precond, "parallelism": "slab%d" % world,
                   "kernels": {k: kernels[k] for k in ("zu", "cg_step", "cg_prec", "collectives")},
                   "l2": "working set per pass (u: %.2f GB) exceeds the 126 MB L2" % (2 * Rl * esz / 1e9)},
        "admm_iters_per_sec": passes / dev_s, "inner_cg_iters_per_pass": J,
        "pass_alg_bytes": b_iter, "pass_gbs": b_iter / (dev_s / passes) / 1e9, "pass_frac_of_peak": b_iter / (dev_s / passes) / 1e9 / peak,
        "pass_kernel_bytes": b_moved, "pass_kernel_gbs": b_moved / (dev_s / passes) / 1e9,
        "pass_kernel_frac_of_peak": b_moved / (dev_s / passes) / 1e9 / peak,
        "wall_seconds": wall, "device_seconds": dev_s, "gen_seconds": t_gen,
        "clocks": clocks, "gpu_launches": launches, "roofline": roof, "stages": stages, "e2e": e2e,
    }
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_baseline(args, threads=0, passes=12, warm=1)   # ~5-10 s of CPU work on the box's 16 host threads
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
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--mode", default="rcpp", choices=["rcpp", "cpp", "py"])
    ap.add_argument("--lam", type=float, default=1.0)
    ap.add_argument("--cg-rtol", dest="cg_rtol", type=float, default=1e-13)
    ap.add_argument("--precond", default="cheb1", choices=["cheb1", "jacobi"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print("bench: warning: fewer than 3 warm-up steps", file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
