"""Probe: host->device copy speed of pageable vs page-locked inputs through Plan.set_points and raw cudart."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multivartv_b200 as mv  # noqa: E402

n = 1 << 24
m = [4096, 4096]
rng = np.random.default_rng(1)
x = rng.random((n, 2))
y = rng.random(n)
axes = [np.linspace(0, 1, d) for d in m]
plan = mv.Plan(m)
xp, yp = mv.pinned_empty(x.shape), mv.pinned_empty(y.shape)
xp[...] = x
yp[...] = y
xf = mv.pinned_empty((2, n)).T          # F-ordered pinned view
xf[...] = x
for name, (a, b) in {"pageable C": (x, y), "pinned C": (xp, yp), "pinned F": (xf, yp), "pinned x, pageable y": (xp, y),
                     "pageable x, pinned y": (x, yp)}.items():
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        plan.set_points(a, b, axes)
        ts.append(time.perf_counter() - t0)
    print("set_points %-22s %s" % (name, ["%.4f" % t for t in ts]), flush=True)

rt = C.CDLL("libcudart.so.12")
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
nb = 8 * n * 2
d = C.c_void_p()
assert rt.cudaMalloc(C.byref(d), nb) == 0
for name, arr in (("pageable", x), ("pinned", xp)):
    for rep in range(3):
        t0 = time.perf_counter()
        assert rt.cudaMemcpy(d, C.c_void_p(arr.ctypes.data), nb, 1) == 0
        rt.cudaDeviceSynchronize()
        dt = time.perf_counter() - t0
        print("cudaMemcpy H2D %-9s %.4f s  %.1f GB/s" % (name, dt, nb / dt / 1e9), flush=True)
out_pg, out_pin = np.empty(n * 2), mv.pinned_empty(n * 2)
for name, arr in (("pageable", out_pg), ("pinned", out_pin)):
    for rep in range(3):
        t0 = time.perf_counter()
        assert rt.cudaMemcpy(C.c_void_p(arr.ctypes.data), d, nb, 2) == 0
        dt = time.perf_counter() - t0
        print("cudaMemcpy D2H %-9s %.4f s  %.1f GB/s" % (name, dt, nb / dt / 1e9), flush=True)
