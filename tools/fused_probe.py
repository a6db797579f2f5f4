"""Probe (round 2): the EXPERIMENTAL degree 2..4 polynomial preconditioner k_cg_horner2d (MVTV_CHEB_DEGREE) and the EXPERIMENTAL fused vector-update + preconditioner kernel k_cg_updprec2d (MVTV_FUSE_UPDPREC=1) against
the separate k_cg_update + k_cg_step2d<STEP_PREC> on 2-D meshes: parity on awkward shapes, then ms/pass on 4096^2.

    python tools/fused_probe.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multivartv_b200 as mv  # noqa: E402
from bench import synth_points  # noqa: E402


def run(m, x, y, axes, env, passes):
    for k in ("MVTV_FUSE_UPDPREC", "MVTV_FUSE_CFG", "MVTV_INIT2D", "MVTV_CHEB_DEGREE", "MVTV_CHEB_KAPPA", "MVTV_HORNER_CFG"):
        os.environ.pop(k, None)
    os.environ.update(env)
    with mv.Plan(m) as plan:
        plan.set_points(x, y, axes)
        kw = dict(mode="rcpp", cg_rtol=1e-13, want_fitted=False, raise_on_nonconvergence=False, precond=mv.PRECOND_CHEB1)
        rw = plan.solve(1.0, max_passes=3, want_theta=False, **kw)
        plan.profile(True)
        r = plan.solve(1.0, max_passes=passes, flags=mv.WARM_THETA_FROM_PLAN | mv.WARM_U_FROM_PLAN, rho_init=rw["rho"],
                       rho_matrix0=rw["rho"], **kw)
        prof = plan.get_profile()
    return r, prof


def main():
    tiny = "--tiny" in sys.argv
    for m, n in (([66, 5], 400), ([2, 9], 60)) if tiny else (([100, 37], 3000), ([66, 5], 400), ([2, 9], 60), ([258, 33], 9000)):
        x, y = synth_points(n, 2, 5)
        axes = [np.linspace(0.0, 1.0, d) for d in m]
        ref, _ = run(m, x, y, axes, {}, 2 if tiny else 12)
        for cfg in range(5):
            env = {"MVTV_INIT2D": "1"} if cfg == 4 else {"MVTV_FUSE_UPDPREC": "1", "MVTV_FUSE_CFG": str(cfg)}   # 4: marching k_cg_init2d alone
            got, _ = run(m, x, y, axes, env, 2 if tiny else 12)
            err = float(np.abs(got["theta"] - ref["theta"]).max())
            print("parity m=%s cfg=%d: max|dtheta|=%.2e passes %d vs %d inner %d vs %d %s" % (
                m, cfg, err, got["passes"], ref["passes"], got["inner_iters"], ref["inner_iters"],
                "ok" if err <= 1e-10 and got["passes"] == ref["passes"] else "MISMATCH"), flush=True)
        for deg in (2, 3, 4):   # k_cg_horner2d: another preconditioner, so the CG path differs; theta agrees to the CG tolerance
            got, _ = run(m, x, y, axes, {"MVTV_CHEB_DEGREE": str(deg)}, 2 if tiny else 12)
            err = float(np.abs(got["theta"] - ref["theta"]).max())
            print("parity m=%s horner degree %d: max|dtheta|=%.2e passes %d vs %d inner %d vs %d %s" % (
                m, deg, err, got["passes"], ref["passes"], got["inner_iters"], ref["inner_iters"],
                "ok" if err <= 1e-9 and got["passes"] == ref["passes"] else "MISMATCH"), flush=True)
    m, n = [4096, 4096], 1 << 24
    if "--tiny" in sys.argv:   # dry run on the CPU emulator (tests/cuda_emu/emu_run.py): only the code path matters
        m, n = [64, 48], 3000
    x, y = synth_points(n, 2, 117)
    axes = [np.linspace(0.0, 1.0, d) for d in m]
    for name, env in [("separate", {}), ("init2d", {"MVTV_INIT2D": "1"})] + [("fused%d" % c, {"MVTV_FUSE_UPDPREC": "1", "MVTV_FUSE_CFG": str(c)}) for c in range(4)] + [
            ("horner%d/k%d" % (d, k), {"MVTV_CHEB_DEGREE": str(d), "MVTV_CHEB_KAPPA": str(k)}) for d in (2, 3, 4) for k in (20, 30)] + [
            ("horner3/k30/c1", {"MVTV_CHEB_DEGREE": "3", "MVTV_HORNER_CFG": "1"})]:
        r, prof = run(m, x, y, axes, env, 2 if "--tiny" in sys.argv else 10)
        inner = r["inner_iters"]
        print("time %-14s ms/pass=%.3f inner/pass=%.1f  us/launch: step=%.1f update(+prec)=%.1f prec=%.1f init=%.1f" % (
            name, 1e3 * r["device_seconds"] / r["passes"], inner / r["passes"], 1e3 * prof["cg_step"][0] / inner,
            1e3 * prof["cg_update"][0] / inner, 1e3 * prof["cg_prec"][0] / max(1, prof["cg_prec"][1]),
            1e3 * prof["cg_init"][0] / max(1, prof["cg_init"][1])), flush=True)


if __name__ == "__main__":
    main()
