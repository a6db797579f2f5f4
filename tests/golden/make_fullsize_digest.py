"""Digest of the CPU oracle's solution on large meshes of the BASELINE families, for the GPU parity tests: the C oracle
(oracle/c, matrix-free Jacobi-PCG to 1e-13) needs minutes per ADMM pass on these meshes, so it is run once here (any CPU
box: the inputs come from bench.py's seeded generator) and a strided sample of theta and u -- every STRIDE-th vertex / row --
is committed together with Counter, rho and the inner-iteration total.  The GPU test
(tests/test_gpu_parity.py::test_large_mesh_digests) solves the same problem with the CUDA path and compares on the sample.
Committed: cfg3_256 (configs[2]'s family on 256^3, n = 2^23) and cfg4_48 (configs[3]'s family on 48^4, n = N).  The full
512^3 / 96^4 meshes (names cfg3, cfg4) need ~60 GB of host memory for the oracle: generate them where that exists and the
same test picks them up.

    python tests/golden/make_fullsize_digest.py cfg3_256 cfg4_48 [cfg3] [cfg4] [--passes 3] [--threads 6]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
STRIDE_THETA, STRIDE_U = 509, 3571     # primes: the samples sweep every axis position


def main():
    from bench import WORKLOADS, synth_points
    from oracle import c_oracle as co
    # the full 512^3 / 96^4 meshes need ~60 GB and hours for the (memory-hungry) oracle; the same families one size down fit any box
    WORKLOADS = dict(WORKLOADS, cfg3_256=dict(m=[256, 256, 256], n=1 << 23), cfg4_48=dict(m=[48, 48, 48, 48], n=48 ** 4))
    passes, threads, names = 3, 6, []
    argv = sys.argv[1:]
    while argv:
        a = argv.pop(0)
        if a == "--passes":
            passes = int(argv.pop(0))
        elif a == "--threads":
            threads = int(argv.pop(0))
        else:
            names.append(a)
    for name in names:
        wl = WORKLOADS[name]
        m, n = wl["m"], wl["n"]
        x, y = synth_points(n, len(m), 117)
        axes = [np.linspace(0.0, 1.0, d) for d in m]
        t0 = time.time()
        ref = co.mbs_one(x, y, m, axes, 1.0, mode=co.MODE_RCPP, max_passes=passes, solver=co.SOLVER_PCG, cg_rtol=1e-13,
                         nthreads=threads)
        dt = time.time() - t0
        out = os.path.join(ROOT, "tests", "golden", "fullsize_%s_digest.npz" % name)
        np.savez_compressed(out, mesh=np.array(m), n=n, passes=passes, lam=1.0, counter=ref["counter"], rho=ref["rho"],
                            inner_iters=ref["inner_iters"], r_norm=ref["r_norm"], s_norm=ref["s_norm"],
                            stride_theta=STRIDE_THETA, stride_u=STRIDE_U,
                            theta=ref["theta"][::STRIDE_THETA].copy(), u=ref["u"][::STRIDE_U].copy(),
                            theta_sum=float(ref["theta"].sum()), theta_abs_max=float(np.abs(ref["theta"]).max()))
        print("%s mesh=%s n=%d passes=%d: Counter %d rho %g inner %d, oracle %.0f s on %d threads -> %s (%d theta, %d u samples)" % (
            name, "x".join(map(str, m)), n, passes, ref["counter"], ref["rho"], ref["inner_iters"], dt, threads, out,
            ref["theta"][::STRIDE_THETA].size, ref["u"][::STRIDE_U].size), flush=True)


if __name__ == "__main__":
    main()
