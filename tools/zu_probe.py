"""Probe (round 2): tile variants of k_zu_march (MVTV_ZU_CFG) on the 3-D and 4-D BASELINE meshes: per-launch time of the z/u
stage and its fraction of the HBM peak (T*(2R+4N) algorithmic bytes), plus parity against the default tile.

    python tools/zu_probe.py [--big]        # --big: 512^3 and 96^4 instead of 256^3 and 48^4
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multivartv_b200 as mv  # noqa: E402
from bench import synth_points  # noqa: E402


def run(m, x, y, axes, cfg, passes=4):
    os.environ["MVTV_ZU_CFG"] = str(cfg)
    with mv.Plan(m) as plan:
        plan.set_points(x, y, axes)
        kw = dict(mode="rcpp", want_fitted=False, raise_on_nonconvergence=False, precond=mv.PRECOND_CHEB1)
        plan.solve(1.0, max_passes=2, want_theta=False, **kw)
        plan.profile(True)
        r = plan.solve(1.0, max_passes=passes, flags=mv.WARM_THETA_FROM_PLAN | mv.WARM_U_FROM_PLAN, **kw)
        prof = plan.get_profile()
        R, N = plan.R, plan.N
    return r, prof, R, N


def main():
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    big = "--big" in sys.argv
    tiny = "--tiny" in sys.argv   # dry run on the CPU emulator (tests/cuda_emu/emu_run.py): only the code path matters
    for m in ([16] * 3, [6] * 4) if tiny else ([512] * 3 if big else [256] * 3, [96] * 4 if big else [48] * 4):
        N = int(np.prod(m))
        x, y = synth_points(N // 2 if len(m) == 3 else N, len(m), 117)
        axes = [np.linspace(0.0, 1.0, d) for d in m]
        ref = None
        for cfg in range(5):
            r, prof, R, Nv = run(m, x, y, axes, cfg)
            ms = prof["zu"][0] / max(1, prof["zu"][1])
            gbs = 8 * (2 * R + 4 * Nv) / (ms * 1e-3) / 1e9
            ref = r if ref is None else ref
            err = float(np.abs(r["theta"] - ref["theta"]).max())
            print("zu %s cfg=%d: %.3f ms/launch, %.0f GB/s = %.2f of peak; ms/pass=%.2f; max|dtheta| vs cfg 0 = %.1e" % (
                "x".join(map(str, m)), cfg, ms, gbs, gbs / peak, 1e3 * r["device_seconds"] / r["passes"], err), flush=True)
    os.environ.pop("MVTV_ZU_CFG", None)


if __name__ == "__main__":
    main()
