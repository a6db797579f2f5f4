#!/usr/bin/env python
"""Golden vectors from the REFERENCE's own compiled C++ solvers (oracle/_ref/libmvtv_ref.so = cpp-code/{utils,solvers}.cpp,
oracle/_ref/libmvtv_ref_rcpp.so = rcpp-code/MultivarTV/src/{utils,solvers}.cpp of /root/reference, built against
oracle/arma_shim/{armadillo,RcppArmadillo.h}, see oracle/ref_shim/Makefile).

    python tests/golden/make_golden_ref_cpp.py      # writes tests/golden/ref_cpp_golden.npz and ref_rcpp_golden.npz

Inputs are regenerated from seeds by tests/helpers.synth, so only the reference's outputs are stored: Counter, theta and
fitted of stand-alone mbs_one solves (CPP mode), a warm-started lambda path on mbs()'s delta-scaled operators, lambda_max,
and dense D for the quirked p = 3 / p = 4 stacks."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_oracle as ro  # noqa: E402
from tests.helpers import synth  # noqa: E402

SOLVES = [(117, 1000, [32, 32], 0.2), (117, 1000, [32, 32], 1.0), (117, 1000, [32, 32], 1.5), (117, 1000, [32, 32], 7.3),
          (5, 300, [16], 2.0), (6, 900, [6, 6, 6], 1.7), (7, 1500, [4, 4, 4, 4], 3.0), (117, 10000, [20, 20], 2.5)]
PATH = (5, 800, [14, 14], np.flipud(np.exp(np.linspace(np.log(6e-3), np.log(6.0), 8))))
LMAX = [(117, 1000, [32, 32]), (117, 10000, [20, 20])]
RSOLVES = [(117, 1000, [32, 32], 0.2), (117, 1000, [32, 32], 1.0), (117, 1000, [32, 32], 1.5), (117, 1000, [32, 32], 7.3),
           (5, 300, [16], 2.0), (6, 900, [6, 6, 6], 1.7), (7, 1500, [4, 4, 4, 4], 3.0)]
RPATH = (5, 800, [14, 14], np.flipud(np.exp(np.linspace(np.log(6e-3), np.log(6.0), 6))))
DMATS = [([3, 3, 3], None), ([3, 3, 3, 3], None), ([4, 4, 4], [0.3, 0.5, 2.0])]

if __name__ == "__main__":
    assert ro.available(), "needs /root/reference (or a prebuilt oracle/_ref/libmvtv_ref.so)"
    out = {"solve_cases": np.array([[s, n, len(m), lam] + m + [0] * (4 - len(m)) for s, n, m, lam in SOLVES], dtype=np.float64)}
    for k, (seed, n, m, lam) in enumerate(SOLVES):
        x, y = synth(seed, n, len(m))
        r = ro.mbs_one(x, y, m, lam)
        out["solve%d_theta" % k], out["solve%d_fitted" % k], out["solve%d_counter" % k] = r["theta"], r["fitted"], np.int64(r["counter"])
    seed, n, m, lambdas = PATH
    x, y = synth(seed, n, len(m))
    r = ro.mbs_path(x, y, m, lambdas=lambdas)
    out["path_case"] = np.array([seed, n] + m, dtype=np.float64)
    out["path_lambdas"], out["path_thetas"], out["path_mses"], out["path_counters"] = lambdas, r["thetas"], r["mses"], r["counters"].astype(np.int64)
    out["lmax_cases"] = np.array([[s, n] + m for s, n, m in LMAX], dtype=np.float64)
    out["lmax_values"] = np.array([ro.lambda_max(*synth(s, n, len(m)), m) for s, n, m in LMAX])
    for dims, deltas in DMATS:
        out["D_" + "x".join(map(str, dims)) + ("_deltas" if deltas else "")] = ro.create_D(dims, deltas)
    path = os.path.join(ROOT, "tests", "golden", "ref_cpp_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")

    # ---- the Rcpp-side sibling: stand-alone mbs_one (theta = mean(y), u = 0, rho = lambda/5, first matrix crossO +
    # lambda*crossD), the warm-started path of mbs_impl (theta, u, rho carried) and mbs_impl with folds = 1
    R = ro.rcpp
    out = {"solve_cases": np.array([[s, n, len(m), lam] + m + [0] * (4 - len(m)) for s, n, m, lam in RSOLVES], dtype=np.float64)}
    for k, (seed, n, m, lam) in enumerate(RSOLVES):
        x, y = synth(seed, n, len(m))
        r = R.mbs_one(x, y, m, lam)
        out["solve%d_theta" % k], out["solve%d_fitted" % k], out["solve%d_u" % k] = r["theta"], r["fitted"], r["u"]
        out["solve%d_counter" % k], out["solve%d_rho" % k] = np.int64(r["counter"]), np.float64(r["rho"])
    seed, n, m, lambdas = RPATH
    x, y = synth(seed, n, len(m))
    r = R.mbs_path(x, y, m, lambdas=lambdas)
    out["path_case"] = np.array([seed, n] + m, dtype=np.float64)
    out["path_lambdas"], out["path_thetas"], out["path_mses"] = lambdas, r["thetas"], r["mses"]
    out["path_counters"], out["path_rhos"] = r["counters"].astype(np.int64), r["rhos"]
    f1 = R.mbs_impl_folds1(x, y, m, lambdas)
    out["impl1_theta"], out["impl1_fitted"], out["impl1_cv_mses"], out["impl1_best"] = f1["theta_hat"], f1["fitted"], f1["cv.mses"], np.int64(f1["lambda_minmse_ind"])
    out["lmax_cases"] = np.array([[s, n] + m for s, n, m in LMAX], dtype=np.float64)
    out["lmax_values"] = np.array([R.lambda_max(*synth(s, n, len(m)), m) for s, n, m in LMAX])
    path = os.path.join(ROOT, "tests", "golden", "ref_rcpp_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")
