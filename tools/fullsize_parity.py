"""BASELINE configs at their stated sizes against the CPU oracle: the CUDA path and oracle/c (PCG to
1e-13, all host threads) run the same `passes` RCPP passes on bench.py's synthetic inputs; identical Counter, theta <= 1e-9,
u <= 1e-8, rho equal.  Minutes of CPU time per case at 512^3 / 96^4, so this is a tool with a committed log
(profiles/r2_fullsize_parity.log), not a pytest case; tests/test_gpu_parity.py holds the 4096^2 / 256^3 / 48^4 cases.

    python tools/fullsize_parity.py [cfg3] [cfg4] [cfg2] [--passes 3]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import multivartv_b200 as mv
    from bench import WORKLOADS, host_threads, synth_points
    from oracle import c_oracle as co
    passes = 3
    names = []
    argv = sys.argv[1:]
    while argv:
        a = argv.pop(0)
        if a == "--passes":
            passes = int(argv.pop(0))
        else:
            names.append(a)
    bad = 0
    for name in names or ["cfg3"]:
        wl = WORKLOADS[name]
        m, n = wl["m"], wl["n"]
        x, y = synth_points(n, len(m), 117)
        axes = [np.linspace(0.0, 1.0, d) for d in m]
        t0 = time.time()
        with mv.Plan(m) as pl:
            pl.set_points(x, y, axes)
            out = pl.solve(1.0, mode="rcpp", max_passes=passes, want_u=True, want_fitted=False, raise_on_nonconvergence=False)
            d = pl.describe()
        t_gpu = time.time() - t0
        t0 = time.time()
        ref = co.mbs_one(x, y, m, axes, 1.0, mode=co.MODE_RCPP, max_passes=passes, solver=co.SOLVER_PCG, cg_rtol=1e-13,
                         nthreads=host_threads())
        t_cpu = time.time() - t0
        et = float(np.abs(out["theta"] - ref["theta"]).max())
        eu = float(np.abs(out["u"] - ref["u"]).max())
        ok = out["counter"] == ref["counter"] and et <= 1e-9 and eu <= 1e-8 and out["rho"] == ref["rho"]
        bad += not ok
        print("%s %s mesh=%s n=%d passes=%d: Counter %d (oracle %d) max|dtheta|=%.2e max|du|=%.2e rho %g (oracle %g) "
              "kernel=%s degree=%d CG %d (oracle Jacobi-PCG %d)  gpu %.1f s, oracle %.1f s on %d threads" % (
                  "OK  " if ok else "FAIL", name, "x".join(map(str, m)), n, passes, out["counter"], ref["counter"], et, eu,
                  out["rho"], ref["rho"], d["cg_step"], d["last_degree"], out["inner_iters"], ref["inner_iters"], t_gpu, t_cpu,
                  host_threads()), flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
