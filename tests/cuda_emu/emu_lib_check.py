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

ENVKEYS = ["MVTV_INIT2D", "MVTV_FUSE_UPDPREC", "MVTV_FUSE_CFG", "MVTV_CHEB_DEGREE", "MVTV_CHEB_KAPPA", "MVTV_HORNER_CFG", "MVTV_STEP3D",
           "MVTV_STEP3D_CFG", "MVTV_ZU_CFG", "MVTV_STEP2D", "MVTV_STEP2D_CFG", "MVTV_STEP2D_PREC_CFG", "MVTV_ZU", "EMU_NSM", "EMU_OCC"]


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

    def run(m, x, y, axes, env, passes, variant=0, lam=0.8, dtype=mv.F64):
        for k in ENVKEYS:
            os.environ.pop(k, None)
        os.environ.update(env)
        with mv.Plan(m, variant=variant, dtype=dtype) as pl:
            d = pl.describe()
            pl.set_points(x, y, axes)
            out = pl.solve(lam, mode="rcpp", max_passes=passes, precond=mv.PRECOND_CHEB1, raise_on_nonconvergence=False, want_u=True)
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

    # ---- 3. every opt-in kernel path (environment variables read at plan creation) against the default path
    def sweep(m, n, seed, passes, envs, variant=0):
        p = len(m)
        x, y = synth(seed, n, p, 0.0, 1.0, 0.5)
        axes = [np.linspace(0, 1, d) for d in m]
        ref, d0 = run(m, x, y, axes, {}, passes, variant)
        orc = co.mbs_one(x, y, m, axes, 0.8, mode=co.MODE_RCPP, max_passes=passes, variant=variant)
        report(np.abs(ref["theta"] - orc["theta"]).max() <= 1e-9, "mesh %s default path (%s) vs oracle after %d passes: max|dtheta| %.2e, %d CG iterations" % (
            m, d0["cg_step"], passes, np.abs(ref["theta"] - orc["theta"]).max(), ref["inner_iters"]))
        for env, want_step, kind in envs:
            out, d = run(m, x, y, axes, env, passes, variant)
            et, eu = float(np.abs(out["theta"] - ref["theta"]).max()), float(np.abs(out["u"] - ref["u"]).max())
            if kind == "same":       # same arithmetic up to summation order: same iteration count
                ok = et <= 1e-12 and eu <= 1e-12 and out["inner_iters"] == ref["inner_iters"]
            elif kind == "fused":    # one launch less per CG iteration
                ok = et <= 1e-12 and eu <= 1e-12 and out["inner_iters"] == ref["inner_iters"] and out["kernel_launches"] < ref["kernel_launches"] - ref["inner_iters"] // 2
            else:                    # another preconditioner: same solution to the CG tolerance, fewer iterations
                ok = et <= 1e-9 and eu <= 1e-9 and out["inner_iters"] < ref["inner_iters"]
            ok = ok and d["cg_step"] == want_step
            report(ok, "  %-75s %-13s max|dtheta| %.1e max|du| %.1e CG %d (default %d) launches %d (%d)" % (
                env, d["cg_step"], et, eu, out["inner_iters"], ref["inner_iters"], out["kernel_launches"], ref["kernel_launches"]))

    s2 = "k_cg_step2d"
    envs2 = [({"MVTV_STEP2D": "smem"}, "k_cg_step", "same"), ({"MVTV_INIT2D": "1"}, s2, "same"), ({"MVTV_FUSE_UPDPREC": "1"}, s2, "fused"),
             ({"MVTV_FUSE_UPDPREC": "1", "MVTV_FUSE_CFG": "2", "MVTV_INIT2D": "1"}, s2, "fused"),
             ({"MVTV_CHEB_DEGREE": "2"}, s2, "prec"), ({"MVTV_CHEB_DEGREE": "3", "MVTV_HORNER_CFG": "1"}, s2, "prec"),
             ({"MVTV_CHEB_DEGREE": "4", "MVTV_CHEB_KAPPA": "100"}, s2, "prec"),
             ({"MVTV_STEP2D_PREC_CFG": "9"}, s2, "same"), ({"MVTV_STEP2D_PREC_CFG": "10"}, s2, "same"), ({"MVTV_STEP2D_PREC_CFG": "11"}, s2, "same"),
             ({"EMU_NSM": "1", "EMU_OCC": "1"}, s2, "same"), ({"EMU_NSM": "16", "EMU_OCC": "4"}, s2, "same")]
    sweep([66, 40], 2500, 1, 3, envs2[:7] if quick else envs2)
    if not quick:
        sweep([130, 33], 3000, 7, 2, [envs2[2], envs2[4], envs2[7]])
    s3, sh = "k_cg_step3d", "k_cg_step3dh"
    envs3 = [({"MVTV_STEP3D": "shfl"}, s3, "same"), ({"MVTV_STEP3D": "hyb"}, sh, "same"),
             ({"MVTV_STEP3D": "shfl", "MVTV_STEP3D_CFG": "1", "MVTV_ZU_CFG": "1"}, s3, "same"),
             ({"MVTV_STEP3D": "hyb", "MVTV_STEP3D_CFG": "1", "MVTV_ZU_CFG": "3"}, sh, "same"),
             ({"MVTV_STEP3D": "shfl", "MVTV_STEP3D_CFG": "6", "MVTV_ZU_CFG": "2"}, s3, "same"),
             ({"MVTV_STEP3D": "shfl", "MVTV_STEP3D_CFG": "5", "MVTV_ZU_CFG": "4"}, s3, "same"),
             ({"MVTV_STEP3D": "hyb", "MVTV_STEP3D_CFG": "5"}, sh, "same"), ({"MVTV_STEP3D": "hyb", "MVTV_STEP3D_CFG": "6", "EMU_NSM": "16"}, sh, "same")]
    sweep([16, 6, 34], 3000, 2, 2, envs3[:2] if quick else envs3, variant=1)   # non-cubic: the intended operator variant
    if not quick:
        sweep([12, 12, 12], 1500, 3, 2, [({"MVTV_STEP3D": "shfl", "MVTV_STEP3D_CFG": str(c)}, s3, "same") for c in (2, 3, 4, 7)] +
              [({"MVTV_STEP3D": "hyb", "MVTV_STEP3D_CFG": str(c)}, sh, "same") for c in (2, 3, 4)])
        sweep([6, 6, 6, 6], 1500, 4, 2, [({"MVTV_ZU_CFG": str(c)}, "k_cg_step", "same") for c in (1, 2, 3, 4)])
    print("emu_lib: %d failure(s)" % FAIL)
    return 1 if FAIL else 0


if __name__ == "__main__":
    sys.exit(main())
