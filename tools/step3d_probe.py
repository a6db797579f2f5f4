"""Probe (round 2, first GPU call): the EXPERIMENTAL shuffle-based k_cg_step3d against the shared-memory k_cg_step on 3-D
meshes -- parity on awkward shapes first, then per-launch kernel times on 256^3 (and 512^3 with --big).

    python tools/step3d_probe.py [--big]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multivartv_b200 as mv  # noqa: E402
from bench import synth_points  # noqa: E402

NCFG = 8          # k_cg_step3d variants (MVTV_STEP3D=shfl)
NCFG_HYB = 7      # k_cg_step3dh variants (MVTV_STEP3D=hyb)


def run(m, x, y, axes, env, passes, precond, dtype=mv.F64):
    for k in ("MVTV_STEP3D", "MVTV_STEP3D_CFG"):
        os.environ.pop(k, None)
    os.environ.update(env)
    variant = mv.VARIANT_REFERENCE if len(set(m)) == 1 else mv.VARIANT_INTENDED   # the reference operator exists on cubic meshes only
    with mv.Plan(m, dtype=dtype, variant=variant) as plan:
        kern = plan.describe()["cg_step"]
        plan.set_points(x, y, axes)
        kw = dict(mode="rcpp", cg_rtol=1e-13 if dtype == mv.F64 else 1e-5, want_fitted=False, raise_on_nonconvergence=False, precond=precond)
        rw = plan.solve(1.0, max_passes=3, want_theta=False, **kw)
        plan.profile(True)
        r = plan.solve(1.0, max_passes=passes, flags=mv.WARM_THETA_FROM_PLAN | mv.WARM_U_FROM_PLAN, rho_init=rw["rho"],
                       rho_matrix0=rw["rho"], **kw)
        prof = plan.get_profile()
    return r, prof, kern


def main():
    bad = 0
    tiny = "--tiny" in sys.argv   # dry run on the CPU emulator (tests/cuda_emu/emu_run.py): two small meshes
    for m, n in (([8, 8, 8], 600), ([66, 3, 4], 500)) if tiny else (([12, 12, 12], 2000), ([66, 5, 7], 1500), ([2, 9, 4], 60), ([130, 33, 6], 9000), ([20, 3, 20], 900)):
        x, y = synth_points(n, 3, 5)
        axes = [np.linspace(0.0, 1.0, d) for d in m]
        for precond in (mv.PRECOND_CHEB1, mv.PRECOND_JACOBI):
            ref, _, k0 = run(m, x, y, axes, {}, 2 if tiny else 12, precond)
            assert k0 == "k_cg_step"
            for kind, cfg in [("shfl", c) for c in range(NCFG)] + [("hyb", c) for c in range(NCFG_HYB)]:
                got, _, k1 = run(m, x, y, axes, {"MVTV_STEP3D": kind, "MVTV_STEP3D_CFG": str(cfg)}, 2 if tiny else 12, precond)
                assert k1 == ("k_cg_step3d" if kind == "shfl" else "k_cg_step3dh")
                err = float(np.abs(got["theta"] - ref["theta"]).max())
                ok = err <= 1e-10 and got["passes"] == ref["passes"]
                bad += not ok
                print("parity m=%s precond=%d %s cfg=%d: max|dtheta|=%.2e inner %d vs %d %s" % (m, precond, kind, cfg, err, got["inner_iters"], ref["inner_iters"], "ok" if ok else "MISMATCH"), flush=True)
    print("parity mismatches:", bad, flush=True)
    sizes = [([256, 256, 256], 256 ** 3)] + ([([512, 512, 512], 1 << 26)] if "--big" in sys.argv else [])
    if "--tiny" in sys.argv:   # dry run on the CPU emulator (tests/cuda_emu/emu_run.py): only the code path matters
        sizes = [([16, 16, 16], 4000)]
    for m, n in sizes:
        x, y = synth_points(n, 3, 117)
        axes = [np.linspace(0.0, 1.0, d) for d in m]
        for name, env in ([("smem", {})] + [("shfl%d" % c, {"MVTV_STEP3D": "shfl", "MVTV_STEP3D_CFG": str(c)}) for c in range(NCFG)]
                          + [("hyb%d" % c, {"MVTV_STEP3D": "hyb", "MVTV_STEP3D_CFG": str(c)}) for c in range(NCFG_HYB)]):
            r, prof, _ = run(m, x, y, axes, env, 2 if "--tiny" in sys.argv else 5, mv.PRECOND_CHEB1)
            inner = r["inner_iters"]
            print("time %s %-6s ms/pass=%.3f inner/pass=%.1f  us/launch: step=%.1f prec=%.1f update=%.1f" % (
                "x".join(map(str, m)), name, 1e3 * r["device_seconds"] / r["passes"], inner / r["passes"],
                1e3 * prof["cg_step"][0] / inner, 1e3 * prof["cg_prec"][0] / inner, 1e3 * prof["cg_update"][0] / inner), flush=True)


if __name__ == "__main__":
    main()
