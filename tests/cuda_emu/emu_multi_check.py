"""TEST INFRASTRUCTURE ONLY -- the multi-GPU path of the library on the CPU emulator: every rank is a host THREAD of this process.

The emulated library of emu_lib_check.py (solver.cu / setup.cu compiled for the SIMT emulator) runs one plan per thread with
rank / world set; CUDA IPC handles carry raw pointers (cuda_emu_rt.h), the .sys loads / stores of the peer-memory protocol are
host atomics (kernels.cuh under MVTV_CUDA_EMU), and NCCL is the in-process stand-in emu_nccl.cpp, found through MVTV_NCCL_LIB.
So the slab partition, the ghost-plane / flag protocol between neighbouring ranks, the rank-ordered peer reductions and the
NCCL fallback (MVTV_COMM=nccl) all execute -- with real concurrency between the ranks -- and the gathered theta is compared
with the single-process C oracle: identical Counter, max|dtheta| <= 1e-9.  A logic check; a wait that never completes ends in
peer_spin's own time-out (NaN results) or in the caller's `timeout`.

    python tests/cuda_emu/emu_multi_check.py <scratch dir> [quick]
"""
import os
import subprocess
import sys
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main():
    scratch = sys.argv[1]
    quick = len(sys.argv) > 2 and sys.argv[2] == "quick"
    import emu_lib_check
    lib = emu_lib_check.build_emulated_library(scratch)
    nccl = os.path.join(scratch, "libnccl_emu.so")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-fPIC", "-shared", os.path.join(HERE, "emu_nccl.cpp"), "-o", nccl, "-lpthread"])
    os.environ["MVTV_NCCL_LIB"] = nccl
    from multivartv_b200 import _lib
    _lib.LIB_PATH = lib          # this process only
    _lib._lib = None
    import multivartv_b200 as mv
    from multivartv_b200 import partition
    from oracle import c_oracle as co
    from tests.helpers import synth
    _lib.load()

    J, C1 = mv.PRECOND_JACOBI, mv.PRECOND_CHEB1
    cases = [  # dims, n, mode, lambda, max_passes, preconditioner, world, comm ("fold": the default peer path, commits folded into the reducing kernels; "peer": MVTV_FOLD_COMMIT=0)
        ([24, 22], 3000, "rcpp", 1.0, 8, C1, 2, "peer"), ([24, 22], 3000, "cpp", 3.0, 0, J, 2, "peer"),
        ([24, 23], 3000, "rcpp", 1.0, 6, C1, 3, "peer"), ([24, 22], 3000, "rcpp", 1.0, 6, C1, 2, "nccl"),
        ([8, 8, 9], 2000, "rcpp", 0.7, 5, C1, 2, "peer"), ([8, 8, 9], 2000, "rcpp", 0.7, 5, J, 3, "peer"),
        ([5, 5, 5, 7], 1500, "rcpp", 1.0, 4, C1, 2, "peer"), ([8, 8, 9], 2000, "cpp", 2.0, 0, C1, 2, "nccl"),
        ([23, 22], 3000, "rcpp", 1.0, 6, C1, 2, "peer"),   # odd width: the shared-memory k_cg_step instead of k_cg_step2d
        ([24, 22], 3000, "rcpp", 1.0, 8, C1, 2, "fold"), ([24, 23], 3000, "cpp", 3.0, 0, J, 3, "fold"),
        ([8, 8, 9], 2000, "rcpp", 0.7, 5, C1, 3, "fold"), ([5, 5, 5, 7], 1500, "rcpp", 1.0, 4, J, 2, "fold"),
        ([23, 22], 3000, "rcpp", 1.0, 6, C1, 2, "fold"),
    ]
    for comm in ("peer", "fold", "nccl"):   # edge cases: one plane per rank, ranks without points, width 1 and 2, world 4
        cases += [([6, 4], 200, "rcpp", 1.0, 4, C1, 4, comm), ([10, 5], 300, "cpp", 2.0, 0, J, 4, comm), ([4, 4, 4], 300, "rcpp", 0.7, 3, C1, 4, comm),
                  ([3, 3, 3, 3], 300, "rcpp", 1.0, 3, J, 3, comm), ([7, 5], 100, "py", 0.8, 4, C1, 2, comm), ([2, 8], 50, "rcpp", 0.3, 4, C1, 4, comm),
                  ([1, 6], 40, "cpp", 1.5, 0, C1, 3, comm), ([16, 7], 5, "rcpp", 1.0, 3, C1, 3, comm)]
    for comm in ("peer", "fold"):          # the 3-D strip kernels on even widths, the ring on every width (9th field: extra environment)
        cases += [([12, 12, 12], 1500, "rcpp", 0.7, 3, C1, 2, comm), ([66, 4, 9], 2000, "rcpp", 1.0, 2, C1, 2, comm),
                  ([8, 8, 8], 600, "rcpp", 0.7, 3, J, 3, comm), ([12, 12, 12], 1500, "rcpp", 0.7, 3, C1, 2, comm, {"MVTV_STEP": "ring"}),
                  ([24, 22], 3000, "rcpp", 1.0, 4, mv.PRECOND_CHEB3, 3, comm), ([24, 22], 3000, "rcpp", 1.0, 4, mv.PRECOND_AUTO, 2, comm)]
    # 3-D strip kernels over the peer path with folded commits: fused update (q exchanged, r kept up to date on the ghost planes)
    # and Horner passes of degree 2..4 (pass outputs rotate over z / idle direction buffer / y, one flag event per exchange)
    cases += [([12, 12, 12], 1500, "rcpp", 0.7, 4, mv.PRECOND_CHEB3, 2, "fold"), ([8, 8, 9], 2000, "rcpp", 0.7, 5, mv.PRECOND_CHEB2, 3, "fold"),
              ([8, 8, 10], 2000, "rcpp", 0.7, 4, mv.PRECOND_CHEB4, 4, "fold"), ([12, 12, 12], 1500, "rcpp", 0.7, 6, mv.PRECOND_AUTO, 3, "fold"),
              ([66, 4, 9], 2000, "rcpp", 1.0, 3, mv.PRECOND_CHEB3, 2, "fold"), ([12, 12, 12], 1500, "cpp", 2.0, 0, mv.PRECOND_CHEB4, 2, "fold")]
    if quick:
        cases = [cases[0], cases[3], cases[4], cases[9], cases[-6], cases[-4]]
    if len(sys.argv) > 2 and sys.argv[2] == "last":
        cases = cases[-6:]
    fail = 0
    for case in cases:
        dims, n, mode, lam, max_passes, precond, world, comm = case[:8]
        extra = case[8] if len(case) > 8 else {}
        p = len(dims)
        imode = {"cpp": 0, "rcpp": 1, "py": 2}[mode]
        x, y = synth(41 + p, n, p, 0.0, 1.0, 0.5)
        axes = mv.mesh_axes(x, dims, mode)
        variant = mv.VARIANT_REFERENCE if (p < 3 or len(set(dims)) == 1) else mv.VARIANT_INTENDED
        buckets = partition.bucket_points(x, y, axes[-1], world)
        os.environ.pop("MVTV_COMM", None)
        os.environ.pop("MVTV_FOLD_COMMIT", None)
        if comm == "nccl":
            os.environ["MVTV_COMM"] = "nccl"
        elif comm == "peer":
            os.environ["MVTV_FOLD_COMMIT"] = "0"   # separate commit launches (the default folds them into the reducing kernels)
        for k in ("MVTV_STEP",):
            os.environ.pop(k, None)
        os.environ.update(extra)
        uid = mv.nccl_unique_id()
        results, errors = [None] * world, []

        def work(rank):
            try:
                with mv.Plan(dims, variant=variant, device=0, rank=rank, world=world, nccl_unique_id=uid) as pl:
                    d = pl.describe()
                    pl.set_points(buckets[rank][0], buckets[rank][1], axes)
                    out = pl.solve(lam, mode=mode, max_passes=max_passes, want_fitted=False, cg_rtol=1e-13, precond=precond,
                                   raise_on_nonconvergence=False)
                    results[rank] = (pl.z0, out["theta"], out["counter"], out["inner_iters"], d, out["kernel_launches"])
            except Exception as e:   # noqa: BLE001
                errors.append((rank, repr(e)))

        threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors or any(r is None for r in results):
            fail += 1
            print("FAIL dims=%s world=%d %s: %s" % (dims, world, comm, errors), flush=True)
            continue
        results.sort(key=lambda t: t[0])
        theta = np.concatenate([r[1] for r in results])
        ref = co.mbs_one(x, y, dims, axes, lam, mode=imode, max_passes=max_passes, variant=variant)
        err = float(np.abs(theta - ref["theta"]).max())
        d = results[0][4]
        good = (all(r[2] == ref["counter"] for r in results) and err <= 1e-9 and d["world"] == world
                and d["collectives"] == ("peer" if comm == "fold" else comm) and d["fold_commit"] == (1 if comm == "fold" else 0))
        fail += not good
        print("%s dims=%s world=%d mode=%s precond=%d collectives=%s%s kernel=%s: Counter %d (oracle %d), %d CG iterations, %d launches, max|dtheta| %.2e" % (
            "ok  " if good else "FAIL", dims, world, mode, precond, comm, (" " + str(extra)) if extra else "", d["cg_step"], results[0][2], ref["counter"], results[0][3], results[0][5], err), flush=True)
    os.environ.pop("MVTV_COMM", None)
    os.environ.pop("MVTV_FOLD_COMMIT", None)
    os.environ.pop("MVTV_STEP", None)
    print("emu_multi: %d failure(s)" % fail)
    return 1 if fail else 0


if __name__ == "__main__":
    sys.exit(main())
