#!/bin/bash
# Multi-GPU call: parity of the slab-partitioned path against the oracle, then strong scaling of cfg3 in a few variants.
#   gpurun --gpus N --timeout 900 -- 'bash tools/gpu_call_multi.sh <tag> N [check] [bench...]'
# BUDGET_S (default 420): wall-clock cap for the WHOLE call -- every step's timeout is cut to what is left, because an N-GPU call is
# charged N x its duration (round 2 lost its last 120 GPU-minutes to a call whose step timeouts added up to more than the limit).
TAG=$1; N=$2; shift 2
OUT=gpurun_out
T0=$(date +%s); BUDGET_S=${BUDGET_S:-420}
left() { local l=$(( BUDGET_S - ($(date +%s) - T0) )); [ $l -lt 5 ] && l=5; [ $l -gt $1 ] && l=$1; echo $l; }
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
PORT=29600
for step in "$@"; do
  PORT=$((PORT+1))
  case $step in
    check) timeout $(left 400) $RUN --master-port $PORT tests/multigpu_check.py > $OUT/${TAG}_n${N}_check.log 2>&1; echo "check rc=$?"; grep -c "^OK" $OUT/${TAG}_n${N}_check.log; grep "FAIL\|all cases\|Error\|error" $OUT/${TAG}_n${N}_check.log | tail -5 ;;
    bench) timeout $(left 300) $RUN --master-port $PORT bench.py --gpus $N --steps 10 --warmup 3 --no-blocks > $OUT/${TAG}_n${N}_bench.json 2> $OUT/${TAG}_n${N}_bench.err; echo "bench rc=$?"; python tools/bench_summary.py $OUT/${TAG}_n${N}_bench.json || tail -5 $OUT/${TAG}_n${N}_bench.err ;;
    bench_cheb1) timeout $(left 300) $RUN --master-port $PORT bench.py --gpus $N --steps 10 --warmup 3 --no-blocks --no-e2e --precond cheb1 > $OUT/${TAG}_n${N}_bench_cheb1.json 2> $OUT/${TAG}_n${N}_bench_cheb1.err; echo "bench_cheb1 rc=$?"; python tools/bench_summary.py $OUT/${TAG}_n${N}_bench_cheb1.json || tail -5 $OUT/${TAG}_n${N}_bench_cheb1.err ;;
    bench_cheb2) timeout $(left 300) $RUN --master-port $PORT bench.py --gpus $N --steps 10 --warmup 3 --no-blocks --no-e2e --precond cheb2 > $OUT/${TAG}_n${N}_bench_cheb2.json 2> $OUT/${TAG}_n${N}_bench_cheb2.err; echo "bench_cheb2 rc=$?"; python tools/bench_summary.py $OUT/${TAG}_n${N}_bench_cheb2.json || tail -5 $OUT/${TAG}_n${N}_bench_cheb2.err ;;
    bench_cheb4) timeout $(left 300) $RUN --master-port $PORT bench.py --gpus $N --steps 10 --warmup 3 --no-blocks --no-e2e --precond cheb4 > $OUT/${TAG}_n${N}_bench_cheb4.json 2> $OUT/${TAG}_n${N}_bench_cheb4.err; echo "bench_cheb4 rc=$?"; python tools/bench_summary.py $OUT/${TAG}_n${N}_bench_cheb4.json || tail -5 $OUT/${TAG}_n${N}_bench_cheb4.err ;;
    bench_unfused) MVTV_TUNE=fused=0 timeout $(left 300) $RUN --master-port $PORT bench.py --gpus $N --steps 10 --warmup 3 --no-blocks --no-e2e --precond cheb1 > $OUT/${TAG}_n${N}_bench_unfused.json 2> $OUT/${TAG}_n${N}_bench_unfused.err; echo "bench_unfused rc=$?"; python tools/bench_summary.py $OUT/${TAG}_n${N}_bench_unfused.json || tail -5 $OUT/${TAG}_n${N}_bench_unfused.err ;;
    bench_r1) MVTV_FOLD_COMMIT=0 timeout $(left 300) $RUN --master-port $PORT bench.py --gpus $N --steps 10 --warmup 3 --no-blocks --no-e2e --precond cheb1 > $OUT/${TAG}_n${N}_bench_r1.json 2> $OUT/${TAG}_n${N}_bench_r1.err; echo "bench_r1 rc=$?"; python tools/bench_summary.py $OUT/${TAG}_n${N}_bench_r1.json || tail -5 $OUT/${TAG}_n${N}_bench_r1.err ;;
    bench_nccl) MVTV_COMM=nccl timeout $(left 300) $RUN --master-port $PORT bench.py --gpus $N --steps 10 --warmup 3 --no-blocks --no-e2e --precond cheb1 > $OUT/${TAG}_n${N}_bench_nccl.json 2> $OUT/${TAG}_n${N}_bench_nccl.err; echo "bench_nccl rc=$?"; python tools/bench_summary.py $OUT/${TAG}_n${N}_bench_nccl.json || tail -5 $OUT/${TAG}_n${N}_bench_nccl.err ;;
    weak2d) timeout $(left 300) $RUN --master-port $PORT bench.py --gpus $N --steps 10 --warmup 3 --no-blocks --no-e2e --workload cfg2 --scaling weak > $OUT/${TAG}_n${N}_weak2d.json 2> $OUT/${TAG}_n${N}_weak2d.err; echo "weak2d rc=$?"; python tools/bench_summary.py $OUT/${TAG}_n${N}_weak2d.json || tail -5 $OUT/${TAG}_n${N}_weak2d.err ;;
  esac
done
