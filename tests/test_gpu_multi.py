"""Launches tests/multigpu_check.py on 2 GPUs and on all GPUs of the box (4 or 8) when they exist (skipped on 1-GPU boxes):
slab partition, peer-memory halos / reductions with the commits folded into the reducing kernels (the default), the
separate-commit variant (MVTV_FOLD_COMMIT=0) and the NCCL path, each against the single-process CPU oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import ctypes as C

    from multivartv_b200 import _lib, build
    build.build()
    n = C.c_int(0)
    _lib.load().mvtv_device_count(C.byref(n))
    return n.value


def _run(world, port, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_check.py")]
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=e)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0
    assert "all cases OK on %d GPUs" % world in r.stdout


def test_two_gpu_slab_parity():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, 29517)
    _run(2, 29518, {"MVTV_FOLD_COMMIT": "0", "MVTV_MG_ONLY2D": "1"})
    _run(2, 29519, {"MVTV_COMM": "nccl", "MVTV_MG_ONLY2D": "1"})


def test_all_gpu_slab_parity():
    """World 4 / 8: interior ranks have both neighbours (the case round 1 only ever ran on the CPU emulator)."""
    n = _ngpu()
    if n < 4:
        pytest.skip("needs >= 4 GPUs")
    _run(8 if n >= 8 else 4, 29520)
