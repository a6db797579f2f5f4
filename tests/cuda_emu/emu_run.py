"""TEST INFRASTRUCTURE ONLY: run a Python script (a probe under tools/, a GPU test module) against the CPU-emulated build of the
library (see emu_lib_check.py) -- a dry run of its host-side logic before GPU minutes are spent on it.  The numbers it prints
(times, GB/s) are meaningless; only "does it run, do the parity lines say ok" counts.

    python tests/cuda_emu/emu_run.py <scratch dir> tools/fused_probe.py --tiny
    python tests/cuda_emu/emu_run.py <scratch dir> -m pytest tests/test_gpu_parity.py -m gpu -k config1
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main():
    scratch = sys.argv[1]
    os.makedirs(scratch, exist_ok=True)
    import emu_lib_check
    lib = emu_lib_check.build_emulated_library(scratch)
    from multivartv_b200 import _lib, build
    _lib.LIB_PATH = lib                      # this process only
    _lib._lib = None
    alias = os.path.join(scratch, "libmvtv_b200.so")
    if not os.path.lexists(alias):
        os.symlink("libmvtv_emu.so", alias)
    build.build = lambda *a, **k: alias      # fixtures call build.build(): nothing to compile with nvcc here
    if sys.argv[2] == "-m":
        sys.argv = [sys.argv[3]] + sys.argv[4:]
        runpy.run_module(sys.argv[0], run_name="__main__", alter_sys=True)
    else:
        sys.argv = sys.argv[2:]
        sys.path.insert(0, os.path.dirname(os.path.abspath(sys.argv[0])))
        runpy.run_path(sys.argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
