"""Probe: shared-memory k_cg_step vs shuffle-based k_cg_step2d on 2-D meshes (per-class kernel times, parity)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multivartv_b200 as mv  # noqa: E402
from bench import synth_points  # noqa: E402


def run(m, x, y, axes, env, passes, precond, dtype=mv.F64):
    for k in ("MVTV_STEP2D", "MVTV_STEP2D_CFG", "MVTV_STEP2D_PREC_CFG"):
        os.environ.pop(k, None)
    os.environ.update(env)
    with mv.Plan(m, dtype=dtype) as plan:
        plan.set_points(x, y, axes)
        kw = dict(mode="rcpp", cg_rtol=1e-13 if dtype == mv.F64 else 1e-5, want_fitted=False, raise_on_nonconvergence=False, precond=precond)
        rw = plan.solve(1.0, max_passes=3, want_theta=False, **kw)
        plan.profile(True)
        r = plan.solve(1.0, max_passes=passes, flags=mv.WARM_THETA_FROM_PLAN | mv.WARM_U_FROM_PLAN, rho_init=rw["rho"],
                       rho_matrix0=rw["rho"], **kw)
        prof = plan.get_profile()
    return r, prof


def main():
    # ---- parity of the two kernels on awkward shapes ---------------------------------------------------------------
    for m, n in (([32, 32], 1000), ([100, 37], 3000), ([66, 5], 400), ([2, 9], 60), ([130, 64], 5000), ([258, 33], 9000)):
        x, y = synth_points(n, 2, 5)
        axes = [np.linspace(0.0, 1.0, d) for d in m]
        for precond in (mv.PRECOND_CHEB1, mv.PRECOND_JACOBI):
            ref, _ = run(m, x, y, axes, {"MVTV_STEP2D": "smem"}, 12, precond)
            for cfg in range(6):
                got, _ = run(m, x, y, axes, {"MVTV_STEP2D": "shfl", "MVTV_STEP2D_CFG": str(cfg % 5), "MVTV_STEP2D_PREC_CFG": str(cfg)}, 12, precond)
                err = float(np.abs(got["theta"] - ref["theta"]).max())
                ok = err <= 1e-10 and got["passes"] == ref["passes"]
                print("parity m=%s precond=%d cfg=%d: max|dtheta|=%.2e inner %d vs %d %s" % (m, precond, cfg, err, got["inner_iters"], ref["inner_iters"], "ok" if ok else "MISMATCH"), flush=True)
        r32, _ = run(m, x, y, axes, {"MVTV_STEP2D": "smem"}, 12, mv.PRECOND_CHEB1, mv.F32)
        g32, _ = run(m, x, y, axes, {"MVTV_STEP2D": "shfl"}, 12, mv.PRECOND_CHEB1, mv.F32)
        print("parity f32 m=%s: max|dtheta|=%.2e" % (m, float(np.abs(g32["theta"] - r32["theta"]).max())), flush=True)

    # ---- timing on BASELINE configs[1] ------------------------------------------------------------------------------
    m, n = [4096, 4096], 1 << 24
    x, y = synth_points(n, 2, 117)
    axes = [np.linspace(0.0, 1.0, d) for d in m]
    for precond, pname in ((mv.PRECOND_CHEB1, "cheb1"), (mv.PRECOND_JACOBI, "jacobi")):
        variants = [("smem", {"MVTV_STEP2D": "smem"})] + [("shfl%d" % c, {"MVTV_STEP2D": "shfl", "MVTV_STEP2D_CFG": str(c % 5), "MVTV_STEP2D_PREC_CFG": str(c)}) for c in (range(6) if pname == "cheb1" else range(5))]
        for name, env in variants:
            r, prof = run(m, x, y, axes, env, 10, precond)
            inner = r["inner_iters"]
            per = {k: (1e3 * v[0] / max(1, inner if k.startswith("cg_") and k != "cg_init" else r["passes"])) for k, v in prof.items() if v[1]}
            print("time %s %-6s ms/pass=%.3f inner/pass=%.1f  us/launch: %s" % (pname, name, 1e3 * r["device_seconds"] / r["passes"], inner / r["passes"],
                  " ".join("%s=%.1f" % (k, v) for k, v in per.items())), flush=True)
    for dtype, dname in ((mv.F32, "f32"),):
        for name, env in (("smem", {"MVTV_STEP2D": "smem"}), ("shfl0", {"MVTV_STEP2D": "shfl"}), ("shfl1", {"MVTV_STEP2D": "shfl", "MVTV_STEP2D_CFG": "1", "MVTV_STEP2D_PREC_CFG": "1"})):
            r, prof = run(m, x, y, axes, env, 10, mv.PRECOND_CHEB1, dtype)
            print("time %s %-6s ms/pass=%.3f inner/pass=%.1f" % (dname, name, 1e3 * r["device_seconds"] / r["passes"], r["inner_iters"] / r["passes"]), flush=True)


if __name__ == "__main__":
    main()
