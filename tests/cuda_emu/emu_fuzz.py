"""TEST INFRASTRUCTURE ONLY: random small cases (p = 1..4, dims 1..14, n = 1..800, all three solver modes, both operator variants,
every preconditioner setting, points all in one cell, delta-scaled operators) through the CPU-emulated library against the C
oracle: identical Counter, max|dtheta| <= 1e-9 (relative to max|theta|).  Not part of the default suite.

    python tests/cuda_emu/emu_fuzz.py <scratch dir> <seed> <cases>

Known non-failures it reports: delta-scaled operators built from one or two distinct points (deltas ~ 2*EPS/m make the system
matrix so ill-conditioned that the oracle's own two solvers, band Cholesky and PCG, disagree by more than the tolerance)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import emu_lib_check  # noqa: E402
from multivartv_b200 import _lib  # noqa: E402

_lib.LIB_PATH = emu_lib_check.build_emulated_library(sys.argv[1])   # this process only
_lib._lib = None
import multivartv_b200 as mv  # noqa: E402
from oracle import c_oracle as co  # noqa: E402

seed0 = int(sys.argv[2])
ncases = int(sys.argv[3])
bad = 0
for case in range(ncases):
    rng = np.random.RandomState(seed0 * 1000 + case)
    p = int(rng.choice([1, 2, 2, 3, 3, 4]))
    variant = int(rng.choice([0, 1]))
    if p >= 3 and variant == 0:
        dims = [int(rng.randint(2, 9 if p == 3 else 6))] * p
    else:
        dims = [int(rng.randint(1, 15 if p <= 2 else (9 if p == 3 else 6))) for _ in range(p)]
    if int(np.prod(dims)) < 2: dims[-1] = 3
    n = int(rng.choice([1, 2, 5, 50, 300, 800]))
    mode = str(rng.choice(["cpp", "rcpp", "py"])); imode = {"cpp": 0, "rcpp": 1, "py": 2}[mode]
    lam = float(rng.choice([0.05, 0.5, 1.0, 2.5, 9.0]))
    precond = int(rng.choice([mv.PRECOND_JACOBI, mv.PRECOND_CHEB1, mv.PRECOND_AUTO]))
    passes = int(rng.choice([1, 2, 5]))
    x = rng.uniform(-1, 1, (n, p)); y = rng.normal(size=n) + (x[:, 0] > 0)
    if rng.rand() < 0.2: x[:] = x[0]          # all points in one cell
    use_d = rng.rand() < 0.3
    try:
        axes = mv.mesh_axes(x, dims, "cpp" if mode == "py" else mode)
        deltas = mv.create_deltas(x, dims, "cpp" if mode == "py" else mode) if use_d else None
        ref = co.mbs_one(x, y, dims, axes, lam, mode=imode, max_passes=passes, variant=variant, deltas=deltas)
        with mv.Plan(dims, deltas=deltas, variant=variant) as pl:
            pl.set_points(x, y, axes)
            out = pl.solve(lam, mode=mode, max_passes=passes, precond=precond, raise_on_nonconvergence=False, cg_rtol=1e-14 if use_d else 0.0)
        err = float(np.abs(out["theta"] - ref["theta"]).max())
        scale = max(1.0, float(np.abs(ref["theta"]).max()))
        ok = out["counter"] == ref["counter"] and err <= 1e-9 * scale
    except Exception as e:
        ok, err = False, repr(e)[:150]
    if not ok:
        bad += 1
    print("%s case %d: p=%d dims=%s variant=%d n=%d mode=%s lam=%g precond=%d passes=%d deltas=%s -> %s" % ("ok  " if ok else "FAIL", seed0 * 1000 + case, p, dims, variant, n, mode, lam, precond, passes, use_d, err if not ok else "%.1e" % err), flush=True)
print("fuzz: %d failure(s)" % bad)
