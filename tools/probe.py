"""Per-kernel-class times of the ADMM pass for a list of solver settings on one mesh (run on a B200 under gpurun).

    python tools/probe.py --mesh 512,512,512 --n 67108864 --passes 6 \
        --variants "auto;cheb1;cheb2;cheb3;cheb4;jacobi;cheb1:fused=0;cheb3:fuse3d=1"

A variant is `precond[:MVTV_TUNE string]` (MVTV_TUNE: the library's developer knob for tile candidates that are still being
measured).  For every variant: max|dtheta| against the first variant after `passes` passes (all variants
solve the same system to cg_rtol, so they must agree to ~1e-10), CG iterations per pass, ms per pass without profiling
events, and the CUDA-event time per launch group of each kernel class."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", default="512,512,512")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--passes", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--lam", type=float, default=1.0)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--variants", default="auto;cheb1;cheb2;cheb3;cheb4;jacobi")
    args = ap.parse_args()
    import multivartv_b200 as mv
    from bench import synth_points
    m = [int(v) for v in args.mesh.split(",")]
    N = int(np.prod(m))
    n = args.n or N
    x, y = synth_points(n, len(m), 117)
    axes = [np.linspace(0.0, 1.0, d) for d in m]
    PRE = {"auto": mv.PRECOND_AUTO, "jacobi": mv.PRECOND_JACOBI, "cheb1": mv.PRECOND_CHEB1, "cheb2": mv.PRECOND_CHEB2,
           "cheb3": mv.PRECOND_CHEB3, "cheb4": mv.PRECOND_CHEB4}
    ref = None
    for var in args.variants.split(";"):
        name, _, tune = var.partition(":")
        os.environ.pop("MVTV_TUNE", None)
        if tune:
            os.environ["MVTV_TUNE"] = tune
        with mv.Plan(m, dtype=mv.F64 if args.dtype == "f64" else mv.F32) as pl:
            pl.set_points(x, y, axes)
            kw = dict(mode="rcpp", cg_rtol=1e-13 if args.dtype == "f64" else 0.0, want_fitted=False, raise_on_nonconvergence=False,
                      precond=PRE[name])
            r = pl.solve(args.lam, max_passes=args.warmup + args.passes, timing_skip_passes=args.warmup, **kw)
            ms_pass = 1e3 * r["device_seconds"] / max(1, r["timed_passes"])
            d = pl.describe()
            pl.profile(True)   # the same call again with the per-class events on (their cost is why it is a second call)
            rp = pl.solve(args.lam, max_passes=args.warmup + 2, want_theta=False, **kw)
            prof = pl.get_profile()
            pl.profile(False)
        if ref is None:
            ref = r["theta"]
        err = float(np.abs(r["theta"] - ref).max())
        inner = rp["inner_iters"]
        perf = {"zu": rp["passes"], "cg_init": rp["passes"], "cg_step": inner, "cg_update": inner}
        deg = d["last_degree"]
        fused = d["fused_update"] and deg >= 1
        perf["cg_prec"] = (rp["passes"] + (inner if deg >= 2 else 0)) if fused else (inner if deg else 0)
        cls = " ".join("%s=%.1f" % (k, 1e3 * prof[k][0] / perf[k]) for k in ("cg_step", "cg_update", "cg_prec", "zu", "cg_init")
                       if perf.get(k) and prof[k][1])
        print("time %-22s mesh=%s degree=%d fused=%d ms/pass=%.3f inner/pass=%.1f max|dtheta|=%.1e  us/launch-group: %s" % (
            var, "x".join(map(str, m)), deg, 1 if fused else 0, ms_pass, r["timed_inner_iters"] / max(1, r["timed_passes"]), err, cls), flush=True)


if __name__ == "__main__":
    main()
