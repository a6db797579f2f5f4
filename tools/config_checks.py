"""BASELINE configs[3] and [4] at full size (run on the GPU box): fp64 vs fp32 on the 4-D 96^4 mesh, and the
32-lambda warm-started path on the 3-D 256^3 mesh.  Prints one JSON line each."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import multivartv_b200 as mv

def synth(n, p, seed):
    rng = np.random.default_rng(seed)
    x = rng.random((n, p))
    y = np.prod(x > 0.5, axis=1) * 1.0 + 0.5 * np.prod(x < 0.2, axis=1) + 0.5 * rng.standard_normal(n)
    return x, y

# configs[3]: 4-D 96^4, fp64 vs fp32
dims = [96] * 4
N = 96 ** 4
x, y = synth(N, 4, 117)
axes = [np.linspace(0, 1, d) for d in dims]
res = {}
for name, dt in (("f64", mv.F64), ("f32", mv.F32)):
    with mv.Plan(dims, dtype=dt) as pl:
        pl.set_points(x, y, axes)
        t = time.time()
        out = pl.solve(1.0, mode="rcpp", max_passes=20, want_fitted=False)
        res[name] = (out["theta"], out["device_seconds"], out["inner_iters"], out["passes"])
d = float(np.abs(res["f64"][0] - res["f32"][0]).max())
print(json.dumps({"check": "configs[3] 4-D 96^4, n=N, RCPP lambda=1, 20 passes: fp32 vs fp64", "max_abs_theta_diff": d,
                  "tolerance": 1e-4, "ok": d <= 1e-4, "f64_ms_per_pass": 1e3 * res["f64"][1] / res["f64"][3],
                  "f32_ms_per_pass": 1e3 * res["f32"][1] / res["f32"][3], "f64_inner": res["f64"][2], "f32_inner": res["f32"][2],
                  "theta_range": [float(res["f64"][0].min()), float(res["f64"][0].max())]}), flush=True)
del res, x, y

# configs[4]: 32 warm-started lambdas on 256^3
dims = [256] * 3
N = 256 ** 3
x, y = synth(N, 3, 118)
axes = [np.linspace(0, 1, d) for d in dims]
with mv.Plan(dims) as pl:
    pl.set_points(x, y, axes)
    lam_max, it = pl.lambda_max("cpp")
    for mode, lams in (("cpp", np.flipud(np.exp(np.linspace(np.log(3e-4), np.log(3.0), 32)))),
                       ("rcpp", np.flipud(np.exp(np.linspace(np.log(3e-2), np.log(3.0), 32))))):
        t = time.time()
        out = pl.solve_path(lams, y, mode=mode, want_best=False, max_counter=(0 if mode == "cpp" else 60))
        wall = time.time() - t
        print(json.dumps({"check": "configs[4] 32-lambda warm-started path, 3-D 256^3, n=N, mode=%s%s" % (mode, "" if mode == "cpp" else " (max_counter=60 per lambda)"),
                          "wall_seconds": wall, "device_seconds": out["device_seconds"], "passes_total": out["passes"],
                          "inner_cg_total": out["inner_iters"], "counters": [int(c) for c in out["counters"]],
                          "vertex_updates_per_sec": N * out["passes"] / out["device_seconds"],
                          "best_index": out["best_index"], "minmse": out["minmse"], "lambda_max_cpp": lam_max, "lambda_max_cg_iters": it}), flush=True)
