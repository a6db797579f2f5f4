"""Probe: STEP_PREC variants of k_cg_step2d (register budget / spill) on 4096^2; parity of each against the default."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multivartv_b200 as mv  # noqa: E402
from bench import synth_points  # noqa: E402
from step2d_probe import run  # noqa: E402,F401  (module-level probe code is guarded below)

CFGS = (0, 6, 9, 10, 11)
TINY = "--tiny" in sys.argv
for m, n in (([66, 6], 300),) if TINY else (([100, 37], 3000), ([258, 33], 9000)):
    x, y = synth_points(n, 2, 5)
    axes = [np.linspace(0.0, 1.0, d) for d in m]
    ref, _ = run(m, x, y, axes, {"MVTV_STEP2D": "smem"}, 2 if TINY else 12, mv.PRECOND_CHEB1)
    for c in CFGS:
        got, _ = run(m, x, y, axes, {"MVTV_STEP2D_PREC_CFG": str(c)}, 2 if TINY else 12, mv.PRECOND_CHEB1)
        err = float(np.abs(got["theta"] - ref["theta"]).max())
        print("parity m=%s prec_cfg=%d: max|dtheta|=%.2e passes %d vs %d %s" % (m, c, err, got["passes"], ref["passes"],
              "ok" if err <= 1e-10 and got["passes"] == ref["passes"] else "MISMATCH"), flush=True)
m, n = [4096, 4096], 1 << 24
if "--tiny" in sys.argv:   # dry run on the CPU emulator (tests/cuda_emu/emu_run.py): only the code path matters
    m, n = [64, 48], 3000
x, y = synth_points(n, 2, 117)
axes = [np.linspace(0.0, 1.0, d) for d in m]
for c in CFGS:
    r, prof = run(m, x, y, axes, {"MVTV_STEP2D_PREC_CFG": str(c)}, 2 if TINY else 10, mv.PRECOND_CHEB1)
    print("time prec_cfg=%d ms/pass=%.3f  us/launch: prec=%.1f step=%.1f update=%.1f" % (c, 1e3 * r["device_seconds"] / r["passes"],
          1e3 * prof["cg_prec"][0] / r["inner_iters"], 1e3 * prof["cg_step"][0] / r["inner_iters"], 1e3 * prof["cg_update"][0] / r["inner_iters"]), flush=True)
