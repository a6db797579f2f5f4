#!/bin/bash
# ncu evidence for bench.py's numbers (run under gpurun, 1 GPU).  Usage: tools/profile.sh <tag> [bench args]
# 1. plain run (must exit 0)  2. launch list with per-launch device time  3. --set full capture of the top kernels
# (recipe: /opt/skills/guides/B200_PROFILING.md; a number printed by a run under ncu is never a bench value)
set -u
TAG=${1:-r2}; shift || true
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-blocks $*"
OUT=gpurun_out
mkdir -p $OUT
python bench.py $ARGS > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py $ARGS > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
# the strip kernels' modes are template arguments: one capture of the family gets them all (-s skips the first launches)
for K in k_cg_step k_zu_march k_cg_init; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 6 -f -o $OUT/${TAG}_$K \
      python bench.py $ARGS > $OUT/${TAG}_ncu_$K.log 2>&1
  echo "$K rc=$?"
  ncu -i $OUT/${TAG}_$K.ncu-rep --page raw --csv > $OUT/${TAG}_${K}_raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_$K.ncu-rep --page details > $OUT/${TAG}_${K}_details.txt 2>/dev/null
done
ls -la $OUT | tail -12
