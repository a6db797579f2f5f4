"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/mvtv.h declares; struct layouts match; compute entry points fail loudly without a GPU (no CPU
fallback).  No compute is performed here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from multivartv_b200 import build
    build.build()
    from multivartv_b200 import _lib
    return _lib.load()


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "mvtv.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mvtv_[a-z_A-Z0-9]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported(lib):
    from multivartv_b200 import _lib
    declared = _declared_symbols()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), "libmvtv_b200.so does not export %s" % name
    assert sorted(_lib.SYMBOLS) == declared
    assert lib.mvtv_abi_version() == 2


def test_struct_layouts_match_header(tmp_path):
    """sizeof/offsetof of the ABI structs as the C compiler sees them == the ctypes mirrors."""
    from multivartv_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "mvtv.h"
int main(void){
  printf("%zu %zu %zu %zu %zu\n", sizeof(mvtv_plan_desc), offsetof(mvtv_plan_desc, dtype), offsetof(mvtv_plan_desc, deltas),
         offsetof(mvtv_plan_desc, nccl_unique_id), offsetof(mvtv_plan_desc, m));
  printf("%zu %zu %zu %zu %zu\n", sizeof(mvtv_solve_params), offsetof(mvtv_solve_params, lambda), offsetof(mvtv_solve_params, max_counter),
         offsetof(mvtv_solve_params, cg_rtol), offsetof(mvtv_solve_params, flags));
  printf("%zu %zu %zu %zu\n", sizeof(mvtv_solve_result), offsetof(mvtv_solve_result, rho), offsetof(mvtv_solve_result, inner_iters),
         offsetof(mvtv_solve_result, kernel_launches));
  return 0; }''')
    exe = tmp_path / "layout"
    subprocess.check_call(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    rows = [list(map(int, ln.split())) for ln in subprocess.check_output([str(exe)]).decode().splitlines()]
    D, P, R = _lib.PlanDesc, _lib.SolveParams, _lib.SolveResult
    assert rows[0] == [C.sizeof(D), D.dtype.offset, D.deltas.offset, D.nccl_unique_id.offset, D.m.offset]
    assert rows[1] == [C.sizeof(P), P.lam.offset, P.max_counter.offset, P.cg_rtol.offset, P.flags.offset]
    assert rows[2] == [C.sizeof(R), R.rho.offset, R.inner_iters.offset, R.kernel_launches.offset]


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product refuses to compute (and says so); it never routes to a CPU path."""
    n = C.c_int(0)
    rc = lib.mvtv_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    import multivartv_b200 as mv
    with pytest.raises(mv.MvtvError) as ei:
        mv.Plan([4, 4])
    assert ei.value.code == 2  # MVTV_ERR_CUDA
    with pytest.raises(mv.MvtvError):
        mv.softthresh(np.ones(3), 0.5)


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/ (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "multivartv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower().replace("no oracle", ""), "%s mentions the oracle" % f


def test_host_mesh_helpers_match_reference_semantics():
    import multivartv_b200 as mv
    rng = np.random.RandomState(3)
    x = rng.uniform(0, 1, (50, 2))
    ax = mv.mesh_axes(x, [4, 5], "cpp")
    assert ax[0][0] == np.float32(x[:, 0].min() + 0.01) and ax[0][-1] == np.float32(x[:, 0].max() + 0.01)
    ax = mv.mesh_axes(x, [4, 5], "rcpp")
    assert ax[1][0] == x[:, 1].min() - 1e-4 and ax[1][-1] == x[:, 1].max() + 1e-4
    mesh = mv.create_mesh(x, [4, 5], "rcpp")
    assert mesh.shape == (20, 2) and mesh[5, 0] == ax[0][1] and mesh[5, 1] == ax[1][1]
    back = mv.axes_from_mesh(mesh, [4, 5])
    assert all(np.array_equal(a, b) for a, b in zip(ax, back))
    d = mv.create_deltas(x, [4, 5], "cpp")
    assert np.allclose(d, (x.max(0) - x.min(0) + 0.02) / np.array([4, 5]))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times beside ours) prints one JSON line with the contract's
    keys; it runs the oracle port on the bounded 128^3 sample of the default workload with ALL host threads -- also under
    torchrun's OMP_NUM_THREADS=1 -- and says in `config` which mesh it really timed; ranks other than 0 print nothing."""
    import json
    import sys
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "mesh_vertex_updates_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["vs_baseline"] is None and d["config"]["mesh"] == [128, 128, 128] and d["config"]["same_config"] is False
    assert d["config"]["sample_of"]["mesh"] == [512, 512, 512] and d["sample_vertices"] == 128 ** 3
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(env, RANK="1", WORLD_SIZE="2"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_bench_byte_accounting_matches_design():
    """bench.py's algorithmic bytes per kernel class add up to the per-pass formulas of DESIGN.md section 3:
    T [(2R+4N) + 8N + 13 N J] with Jacobi, T [(2R+4N) + 8N + (3+4(d-1)) N + (13+4(d-1)) N J] with the fused degree-d polynomial."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    N, R, J = 1000, 6980, 40
    strip = {"cg_prec_words": 3, "fused_update": 1}
    tot, perf = bench.stage_totals(N, R, 8, 1, J, 0, strip)
    assert sum(tot.values()) == 8 * ((2 * R + 4 * N) + 8 * N + 13 * N * J) and perf["cg_prec"] == 0
    for d in (1, 2, 3, 4):
        tot, perf = bench.stage_totals(N, R, 8, 1, J, d, strip)
        assert sum(tot.values()) == 8 * ((2 * R + 4 * N) + 8 * N + (3 + 4 * (d - 1)) * N + (13 + 4 * (d - 1)) * N * J)
        assert perf["cg_step"] == J and perf["cg_update"] == J and perf["cg_prec"] == 1 + (J if d >= 2 else 0)
    ring = {"cg_prec_words": 4, "fused_update": 0}
    tot, _ = bench.stage_totals(N, R, 8, 1, J, 1, ring)          # ring kernels: 5 N + 6 N + 4 N per iteration
    assert sum(tot.values()) == 8 * ((2 * R + 4 * N) + 8 * N + 15 * N * J)
    f = bench.pass_fractions(N, R, 8, J, 1.0, 6545.0, sum(tot.values()))
    assert f["alg_bytes"] == 8 * ((2 * R + 3 * N) + 12 * N * J) and abs(f["frac_of_8TBs"] * 8000.0 - f["gbs"]) < 1e-9
