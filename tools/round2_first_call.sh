#!/bin/bash
# First GPU call of the next round: time (and parity-check) everything that was written after round 1's GPU minutes ran out.
# All of it is opt-in and logic-checked on the CPU emulator (tests/cuda_emu); nothing here changes a default.  Every script below was
# dry-run against the emulated library first (python tests/cuda_emu/emu_run.py <scratch> tools/<probe>.py --tiny), so what can
# still go wrong on the GPU is the kernels' behaviour on real hardware, not the scripts.  The multi-GPU counterpart is
# tools/round2_multi_gpu_call.sh (folded commits, MVTV_FOLD_COMMIT=1).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/round2_first_call.sh'
OUT=gpurun_out
mkdir -p $OUT
timeout 480 python tools/step3d_probe.py --big      > $OUT/r2_step3d_probe.log 2>&1; echo "step3d rc=$?"; grep -c " ok$" $OUT/r2_step3d_probe.log; grep "^time\|MISMATCH\|mismatches" $OUT/r2_step3d_probe.log | tail -40
timeout 200 python tools/zu_probe.py --big          > $OUT/r2_zu_probe.log 2>&1;     echo "zu rc=$?";     tail -12 $OUT/r2_zu_probe.log
timeout 300 python tools/fused_probe.py             > $OUT/r2_fused_probe.log 2>&1;  echo "fused rc=$?";  grep "^time\|MISMATCH" $OUT/r2_fused_probe.log
timeout 120 python tools/step2d_probe_prec.py       > $OUT/r2_prec_probe.log 2>&1;   echo "prec rc=$?";   grep "^time\|MISMATCH" $OUT/r2_prec_probe.log
MVTV_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "step3d or step2d" > $OUT/r2_pytest_experimental.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r2_pytest_experimental.log
