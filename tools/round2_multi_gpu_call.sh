#!/bin/bash
# Second GPU call of the next round (N GPUs of one box, default 2): the EXPERIMENTAL folded commits of the peer-memory CG loop
# (MVTV_FOLD_COMMIT=1: the reducing kernel's last thread also waits for the world's partial sums and commits the CG scalars, so
# the three one-thread k_cg_peer_commit_* launches per CG iteration disappear).  Logic-checked with every rank a host thread on
# the CPU emulator (tests/cuda_emu/emu_multi_check.py); this is its first run on GPUs.  Every command runs under `timeout`.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/round2_multi_gpu_call.sh 2'
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
MVTV_FOLD_COMMIT=1 timeout 300 $RUN --master-port 29541 tests/multigpu_check.py > $OUT/r2_multi_fold_parity.log 2>&1; echo "fold parity rc=$?"; grep -c "^OK" $OUT/r2_multi_fold_parity.log; grep "FAIL\|all cases" $OUT/r2_multi_fold_parity.log | tail -5
for fold in 0 1; do
  MVTV_FOLD_COMMIT=$fold timeout 200 $RUN --master-port 2955$fold bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --workload cfg3 > $OUT/r2_strong_cfg3_n${N}_fold$fold.json 2> $OUT/r2_strong_cfg3_n${N}_fold$fold.err; echo "strong cfg3 fold=$fold rc=$?"; python tools/bench_summary.py $OUT/r2_strong_cfg3_n${N}_fold$fold.json 2>/dev/null || tail -c 400 $OUT/r2_strong_cfg3_n${N}_fold$fold.json
  MVTV_FOLD_COMMIT=$fold timeout 200 $RUN --master-port 2956$fold bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2_weak_cfg2_n${N}_fold$fold.json 2> $OUT/r2_weak_cfg2_n${N}_fold$fold.err; echo "weak cfg2 fold=$fold rc=$?"; python tools/bench_summary.py $OUT/r2_weak_cfg2_n${N}_fold$fold.json 2>/dev/null || tail -c 400 $OUT/r2_weak_cfg2_n${N}_fold$fold.json
done
