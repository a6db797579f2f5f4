"""TEST INFRASTRUCTURE ONLY -- the WHOLE library on the CPU SIMT emulator.

csrc/solver.cu and csrc/setup.cu, with their <<<>>> launches rewritten by emu_translate.py and the CUDA runtime replaced by
the stand-ins of cuda_emu_rt.h, are compiled by g++ into a scratch libmvtv_emu.so; this script (run as a subprocess by
tests/test_cuda_emu.py) points the ctypes binding of THIS process at that scratch library and drives the real Python mirror
over it.  What it checks is the host code no CPU test could reach before: plan set-up, kernel selection by mesh / environment,
chunking, the ADMM and CG drivers, the lambda path, the C ABI -- against the C oracle, and every opt-in kernel path
(MVTV_* environment variables) against the default path.  It is a logic check: it says nothing about performance, the
memory model or PTX, and the product never loads this library (multivartv_b200/_lib.py has one fixed path and no fallback).

    python tests/cuda_emu/emu_lib_check.py <scratch dir> [quick]
"""
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

ENVKEYS = ["MVTV_STEP", "MVTV_ZU_KERNEL", "MVTV_TUNE", "EMU_NSM", "EMU_OCC"]


def build_emulated_library(scratch):
    """g++ build of the translated sources; returns the path of libmvtv_emu.so."""
    from emu_translate import translate
    csrc = os.path.join(ROOT, "multivartv_b200", "csrc")
    lib = os.path.join(scratch, "libmvtv_emu.so")
    newest = max(os.path.getmtime(os.path.join(d, f)) for d in (csrc, HERE, os.path.join(HERE, "fake"), os.path.join(ROOT, "include"))
                 for f in os.listdir(d) if os.path.isfile(os.path.join(d, f)))
    if os.path.exists(lib) and os.path.getmtime(lib) >= newest and os.environ.get("EMU_REBUILD") != "1":
        return lib   # a scratch directory shared by several checks is built once
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    objs, procs = [], []
    for name in ("solver", "setup"):
        text, n = translate(open(os.path.join(csrc, name + ".cu")).read())
        assert n > 0
        cpp = os.path.join(scratch, name + "_emu.cpp")
        open(cpp, "w").write(text)
        obj = os.path.join(scratch, name + "_emu.o")
        objs.append(obj)
        # fibers switch stacks with _longjmp: glibc's fortified longjmp would reject that
        procs.append(subprocess.Popen([gxx, "-std=c++17", "-O1", "-w", "-fPIC", "-U_FORTIFY_SOURCE", "-D_FORTIFY_SOURCE=0", "-DMVTV_BUILD",
                                       "-I", os.path.join(HERE, "fake"), "-I", HERE, "-I", csrc, "-c", cpp, "-o", obj]))
    for p in procs:
        assert p.wait() == 0
    subprocess.check_call([gxx, "-shared", "-o", lib] + objs + ["-ldl"])
    alias = os.path.join(scratch, "libmvtv_b200.so")   # scratch-only alias so that `-L <scratch> -lmvtv_b200` (tests/cpp) links to the emulation
    if not os.path.lexists(alias):
        os.symlink("libmvtv_emu.so", alias)
    return lib


FAIL = 0


def report(ok, text):
    global FAIL
    if not ok:
        FAIL += 1
    print(("ok   " if ok else "FAIL ") + text, flush=True)


def main():
    scratch = sys.argv[1]
    quick = len(sys.argv) > 2 and sys.argv[2] == "quick"
    lib = build_emulated_library(scratch)
    from multivartv_b200 import _lib
    _lib.LIB_PATH = lib          # this process only: the emulated library instead of the CUDA one
    _lib._lib = None
    import multivartv_b200 as mv
    from oracle import c_oracle as co
    from tests.helpers import synth

    def run(m, x, y, axes, env, passes, variant=0, lam=0.8, dtype=mv.F64, precond=mv.PRECOND_CHEB1):
        for k in ENVKEYS:
            os.environ.pop(k, None)
        os.environ.update(env)
        with mv.Plan(m, variant=variant, dtype=dtype) as pl:
            d = pl.describe()
            pl.set_points(x, y, axes)
            out = pl.solve(lam, mode="rcpp", max_passes=passes, precond=precond, raise_on_nonconvergence=False, want_u=True)
            d = pl.describe()
        for k in ENVKEYS:
            os.environ.pop(k, None)
        return out, d

    # ---- 1. the default path against the C oracle: both C++ solver modes to convergence on a small mesh, Counter and theta
    x, y = synth(117, 300, 2)
    m = [16, 12]
    for mode, imode, lam, cap in (("cpp", co.MODE_CPP, 1.5, 0), ("rcpp", co.MODE_RCPP, 1.0, 25 if quick else 0)):
        axes = mv.mesh_axes(x, m, mode)
        ref = co.mbs_one(x, y, m, axes, lam, mode=imode, max_passes=cap)
        t0 = time.time()
        with mv.Plan(m) as pl:
            pl.set_points(x, y, axes)
            out = pl.solve(lam, mode=mode, max_passes=cap, raise_on_nonconvergence=False)
        err = float(np.abs(out["theta"] - ref["theta"]).max())
        report(out["counter"] == ref["counter"] and err <= 1e-9 and out["kernel_launches"] > 0,
               "mbs_one %s mode on %s vs oracle: Counter %d / %d, max|dtheta| %.2e, %d launches (%.1f s)" % (
                   mode, m, out["counter"], ref["counter"], err, out["kernel_launches"], time.time() - t0))

    # ---- 2. lambda path, lambda_max, predict, operators through the ABI
    m = [10, 9]
    x, y = synth(5, 200, 2)
    axes = mv.mesh_axes(x, m, "cpp")
    with mv.Plan(m) as pl:
        pl.set_points(x, y, axes)
        lmax, lmax_it = pl.lambda_max("cpp")
        report(np.isfinite(lmax) and lmax > 0 and lmax_it > 0, "lambda_max (cpp) = %.6g after %d CG iterations" % (lmax, lmax_it))
        lams = np.array([1.5, 0.7, 0.2])
        path = pl.solve_path(lams, y, mode="cpp", want_thetas=True)
        th = np.random.RandomState(0).normal(size=90)
        Dth = pl.apply_D(th)
        op = co.Operator(m)
        report(np.abs(Dth - op.D(th)).max() <= 1e-13, "apply_D vs oracle on %s: %.2e" % (m, np.abs(Dth - op.D(th)).max()))
        w = np.random.RandomState(1).normal(size=Dth.size)
        report(np.abs(pl.apply_Dt(w) - op.Dt(w)).max() <= 1e-12, "apply_Dt vs oracle")
        fit = pl.predict(x, theta=path["thetas"][0])
        report(np.array_equal(fit, path["thetas"][0][co.nearest(m, axes, x)]), "predict = theta[nearest vertex] (bit-exact)")
        theta = None
        ok = True
        for i, lam in enumerate(lams):
            r = co.mbs_one(x, y, m, axes, float(lam), mode=co.MODE_CPP, theta_init=theta)
            theta = r["theta"]
            ok = ok and int(path["counters"][i]) == r["counter"] and np.abs(path["thetas"][i] - theta).max() <= 1e-9
        report(ok, "solve_path (cpp, 3 warm-started lambdas) vs oracle: counters %s" % list(path["counters"]))

    # ---- 3. every kernel family / preconditioner degree the host code can select, against the oracle and against each other:
    # the strip kernels (default on even m0) with the fused update, Jacobi and Horner degrees 1..4, the shared-memory ring
    # (MVTV_STEP=ring), the gather z/u kernel, and other SM counts / occupancies (the chunking of the marching kernels)
    P = {"jacobi": mv.PRECOND_JACOBI, "cheb1": mv.PRECOND_CHEB1, "cheb2": mv.PRECOND_CHEB2, "cheb3": mv.PRECOND_CHEB3,
         "cheb4": mv.PRECOND_CHEB4, "auto": mv.PRECOND_AUTO}

    def sweep(m, n, seed, passes, cases, variant=0):
        p = len(m)
        x, y = synth(seed, n, p, 0.0, 1.0, 0.5)
        axes = [np.linspace(0, 1, d) for d in m]
        orc = co.mbs_one(x, y, m, axes, 0.8, mode=co.MODE_RCPP, max_passes=passes, variant=variant)
        base = {}
        for env, prec, want_step, want_deg in cases:
            out, d = run(m, x, y, axes, env, passes, variant, precond=P[prec])
            et = float(np.abs(out["theta"] - orc["theta"]).max())
            eu = float(np.abs(out["u"] - orc["u"]).max()) if orc.get("u") is not None else 0.0
            ok = et <= 1e-9 and eu <= 1e-8 and out["passes"] == orc["passes"] and d["cg_step"] == want_step and d["last_degree"] == want_deg
            key = (d["last_degree"],)
            if key in base:      # same polynomial degree through another kernel family / launch shape: same iteration count
                ok = ok and out["inner_iters"] == base[key]
            else:
                base[key] = out["inner_iters"]
            report(ok, "  mesh %s %-22s %-7s %-12s degree %d: max|dtheta| %.1e max|du| %.1e CG %d launches %d" % (
                m, env, prec, d["cg_step"], d["last_degree"], et, eu, out["inner_iters"], out["kernel_launches"]))
        its = [base[(k,)] for k in range(5) if (k,) in base]
        report(all(a > b for a, b in zip(its[:2], its[1:2])), "  mesh %s: CG iterations by degree %s" % (m, its))

    s2, s3, ring = "k_cg_step2d", "k_cg_step3d", "k_cg_step"
    R = {"MVTV_STEP": "ring"}
    cases2 = [({}, "jacobi", s2, 0), ({}, "cheb1", s2, 1), ({}, "cheb2", s2, 2), ({}, "cheb3", s2, 3), ({}, "cheb4", s2, 4), ({}, "auto", s2, 3),
              (R, "jacobi", ring, 0), (R, "cheb1", ring, 1), (R, "cheb3", ring, 1),
              ({"EMU_NSM": "1", "EMU_OCC": "1"}, "cheb3", s2, 3), ({"EMU_NSM": "16", "EMU_OCC": "4"}, "cheb2", s2, 2),
              ({"MVTV_ZU_KERNEL": "gather"}, "cheb1", s2, 1)]
    sweep([66, 40], 2500, 1, 3, cases2[:8] if quick else cases2)
    if not quick:
        sweep([130, 33], 3000, 7, 2, [cases2[1], cases2[3], cases2[7]])
    cases3 = [({}, "jacobi", s3, 0), ({}, "cheb1", s3, 1), ({}, "cheb2", s3, 2), ({}, "cheb3", s3, 3), ({}, "cheb4", s3, 4),
              (R, "jacobi", ring, 0), (R, "cheb1", ring, 1), ({"EMU_NSM": "16"}, "cheb3", s3, 3), ({"EMU_NSM": "2", "EMU_OCC": "1"}, "cheb2", s3, 2),
              ({"MVTV_TUNE": "init3d=0"}, "jacobi", s3, 0), ({"MVTV_TUNE": "fused=0"}, "cheb1", s3, 1), ({"MVTV_TUNE": "init3d=0"}, "cheb3", s3, 3),
              ({"MVTV_TUNE": "fused=0"}, "cheb2", s3, 2), ({"MVTV_TUNE": "init3d=0,fused=0", "EMU_NSM": "3"}, "cheb4", s3, 4)]
    sweep([16, 6, 34], 3000, 2, 2, (cases3[:5] + cases3[9:12]) if quick else cases3, variant=1)   # non-cubic: the intended operator variant
    if not quick:
        sweep([12, 12, 12], 1500, 3, 2, [cases3[1], cases3[3], cases3[6], cases3[7]])
        sweep([66, 5, 7], 1500, 5, 2, [cases3[0], cases3[1], cases3[2], cases3[3]] + cases3[9:], variant=1)
        sweep([130, 9, 8], 3000, 6, 2, [cases3[1], cases3[3]] + cases3[9:], variant=1)
        sweep([6, 6, 6, 6], 1500, 4, 2, [({}, "jacobi", ring, 0), ({}, "cheb1", ring, 1), ({}, "cheb4", ring, 1), ({"MVTV_ZU_KERNEL": "gather"}, "cheb1", ring, 1)])
    print("emu_lib: %d failure(s)" % FAIL)
    return 1 if FAIL else 0


if __name__ == "__main__":
    sys.exit(main())
