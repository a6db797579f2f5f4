#!/bin/bash
# ncu evidence for bench.py's numbers (run under gpurun, 1 GPU).  Usage: tools/profile.sh <tag> [bench args]
# 1. plain run (must exit 0)  2. launch list with per-launch device time  3. --set full capture of the top kernels
set -u
TAG=${1:-r1}; shift || true
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e $*"
OUT=gpurun_out
python bench.py $ARGS > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py $ARGS > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
for K in k_cg_step k_cg_update k_zu k_cg_init; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 2 -f -o $OUT/${TAG}_$K \
      python bench.py $ARGS > $OUT/${TAG}_ncu_$K.log 2>&1
  echo "$K rc=$?"
done
ls -la $OUT | tail -20
