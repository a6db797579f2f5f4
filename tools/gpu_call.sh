#!/bin/bash
# One GPU call: tests, probes, bench.  usage: gpurun --timeout 1500 -- 'bash tools/gpu_call.sh <tag> [steps...]'
TAG=${1:-r2}; shift
OUT=gpurun_out
mkdir -p $OUT
for step in "$@"; do
  case $step in
    pytest) timeout 900 python -m pytest tests -m gpu -x -q --durations=12 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/${TAG}_pytest.log ;;
    probe3) timeout 400 python tools/probe.py --mesh 512,512,512 --n 67108864 --passes 4 --variants "auto;cheb1;cheb2;cheb3;cheb4;jacobi;cheb1:fused=0;cheb3:fused=0;cheb1:fuse3d=1;cheb1:fuse3d=2;cheb3:horner3d=1;cheb3:horner3d=2" > $OUT/${TAG}_probe3.log 2>&1; echo "probe3 rc=$?"; grep "^time" $OUT/${TAG}_probe3.log ;;
    probe2) timeout 300 python tools/probe.py --mesh 4096,4096 --n 16777216 --passes 10 --variants "auto;cheb1;cheb2;cheb3;cheb4;jacobi;cheb1:fused=0;cheb3:fused=0" > $OUT/${TAG}_probe2.log 2>&1; echo "probe2 rc=$?"; grep "^time" $OUT/${TAG}_probe2.log ;;
    probe4) timeout 300 python tools/probe.py --mesh 96,96,96,96 --passes 3 --variants "auto;jacobi" > $OUT/${TAG}_probe4.log 2>&1; echo "probe4 rc=$?"; grep "^time" $OUT/${TAG}_probe4.log ;;
    bench) timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; python tools/bench_summary.py $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err ;;
    benchref) timeout 300 python bench.py --impl reference > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "benchref rc=$?"; cut -c1-600 $OUT/${TAG}_bench_ref.json ;;
    full3) timeout 900 python tools/fullsize_parity.py cfg3 --passes 3 > $OUT/${TAG}_fullsize_cfg3.log 2>&1; echo "full3 rc=$?"; tail -2 $OUT/${TAG}_fullsize_cfg3.log ;;
    full4) timeout 900 python tools/fullsize_parity.py cfg4 --passes 3 > $OUT/${TAG}_fullsize_cfg4.log 2>&1; echo "full4 rc=$?"; tail -2 $OUT/${TAG}_fullsize_cfg4.log ;;
    smoke) timeout 200 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/${TAG}_smoke.log ;;
    hostinfo) nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total --format=csv ;;
  esac
done
