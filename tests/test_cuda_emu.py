"""The library's CUDA kernel SOURCE run on a CPU SIMT emulator (tests/cuda_emu/cuda_emu.h: one ucontext fiber per CUDA thread,
barriers for __syncthreads and warp shuffles): logic checks that need no GPU.  tests/cuda_emu/emu_cg_step.cpp compares
k_cg_step, k_cg_step2d and k_cg_step3d with a host loop over the clamped stencil, single rank and slab
by slab.  CPU test; says nothing about performance or PTX-level behaviour."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "cuda_emu")


def _build(tmp_path, name="emu_cg_step"):
    exe = str(tmp_path / name)
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    # fibers switch stacks with _longjmp: glibc's fortified longjmp would reject that
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-w", "-U_FORTIFY_SOURCE", "-D_FORTIFY_SOURCE=0", "-I", os.path.join(EMU, "fake"), "-I", EMU,
                           os.path.join(EMU, name + ".cpp"), "-o", exe])
    return exe


def test_zu_kernels_on_the_emulator(tmp_path):
    """k_zu (gather) vs k_zu_march (scatter, marching) in the tile shapes of zu_march.cuh, 2-D / 3-D / 4-D."""
    exe = _build(tmp_path, "emu_zu")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "emu_zu: 0 failure(s)" in r.stdout


def test_cg_step_kernels_on_the_emulator(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "emu_cg_step: 0 failure(s)" in r.stdout
    # self-test of the comparisons: a 1e-6 perturbation of the kernels' matrix scalar must be caught
    env = dict(os.environ, EMU_NEGATIVE="1")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 1 and "FAIL" in r.stdout


@pytest.fixture(scope="module")
def emu_scratch(tmp_path_factory):
    """One scratch directory for the emulated library: built once (g++, ~1 min), shared by the checks below."""
    return str(tmp_path_factory.mktemp("emu_lib"))


def test_whole_library_on_the_emulator(emu_scratch):
    """tests/cuda_emu/emu_lib_check.py: csrc/solver.cu + setup.cu themselves (launches rewritten by emu_translate.py, CUDA runtime
    replaced by cuda_emu_rt.h) built into a scratch library and driven through the Python mirror in a subprocess: the default
    path against the C oracle (Counter, theta, lambda path, operators) and every kernel family / preconditioner degree (strip
    kernels with the fused update and Horner degrees 1..4, the shared-memory ring) against the oracle and each other, all
    through the real host code (plan set-up, kernel selection, chunking, CG / ADMM drivers, C ABI).  MVTV_EMU_FULL=1 runs the
    full set (~3 min) instead of the reduced one (~1.5 min, most of it the g++ build)."""
    args = [sys.executable, os.path.join(EMU, "emu_lib_check.py"), emu_scratch]
    if os.environ.get("MVTV_EMU_FULL") != "1":
        args.append("quick")
    r = subprocess.run(args, capture_output=True, text=True, timeout=3000)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-3000:]
    assert "emu_lib: 0 failure(s)" in r.stdout and "FAIL" not in r.stdout


def test_multi_rank_path_on_the_emulator(emu_scratch):
    """tests/cuda_emu/emu_multi_check.py: the slab-partitioned multi-GPU solve with every rank a host thread of one process --
    peer-memory ghost planes / flags / rank-ordered reductions running concurrently, and the NCCL fallback over an in-process
    NCCL stand-in (emu_nccl.cpp) -- world 2 and 3, 2-D / 3-D / 4-D, both preconditioners, against the single-process C oracle:
    identical Counter, max|dtheta| <= 1e-9."""
    r = subprocess.run([sys.executable, os.path.join(EMU, "emu_multi_check.py"), emu_scratch], capture_output=True, text=True, timeout=3000)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-3000:]
    assert "emu_multi: 0 failure(s)" in r.stdout and "FAIL" not in r.stdout


def test_launch_translation():
    """emu_translate.py on the launch forms the library uses (template arguments, nested dim3, '->' in the configuration)."""
    sys.path.insert(0, EMU)
    from emu_translate import translate
    src = ("if (a) k_x<T, Cfg<2, 3>, 1><<<dim3((unsigned)t, n, 1), Cfg::NT, smem, plan->stream>>>(dt, RedBuf{p, c + 5}, f(z));\n"
           "else k_y<<<plan->grid(), 256>>>(a,\n   b);\n")
    out, n = translate(src)
    assert n == 2
    assert "::cuda_emu::launch(dim3(dim3((unsigned)t, n, 1)), dim3(Cfg::NT), (size_t)(smem), [&] { k_x<T, Cfg<2, 3>, 1>(dt, RedBuf{p, c + 5}, f(z)); });" in out
    assert "else ::cuda_emu::launch(dim3(plan->grid()), dim3(256), (size_t)(0), [&] { k_y(a,\n   b); });" in out
