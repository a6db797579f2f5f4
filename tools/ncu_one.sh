#!/bin/bash
# usage: tools/ncu_one.sh <tag> <kernel-regex> <skip> <count> [bench args]
TAG=$1; K=$2; SKIP=$3; CNT=$4; shift 4
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e $*"
python bench.py $ARGS > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo plain failed; tail -3 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -f -o gpurun_out/${TAG} python bench.py $ARGS > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}.ncu-rep --page details > gpurun_out/${TAG}_details.txt 2>&1
