#!/bin/bash
# Round-end evidence on 1 GPU: default bench line (+ reference arm), other configs, fp32, ncu launch list + full
# captures (exported to CSV on the box; the .ncu-rep files are too big to travel back).
OUT=gpurun_out
export_rep() {  # <name>
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
  ncu -i $OUT/$1.ncu-rep --page details > $OUT/$1_details.txt 2>/dev/null
  ncu -i $OUT/$1.ncu-rep --page source --csv > $OUT/$1_source.csv 2>/dev/null
  rm -f $OUT/$1.ncu-rep
}
if [ "$1" != "ncu-only" ]; then
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/f_bench_reference.json 2> $OUT/f_bench_reference.err; echo "reference rc=$?"
timeout 400 python bench.py > $OUT/f_bench_default.json 2> $OUT/f_bench_default.err; echo "default rc=$?"
for W in cfg3 cfg4; do timeout 400 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > $OUT/f_bench_$W.json 2> $OUT/f_bench_$W.err; echo "$W rc=$?"; done
timeout 300 python bench.py --dtype f32 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/f_bench_cfg2_f32.json 2> $OUT/f_bench_cfg2_f32.err; echo "f32 rc=$?"
timeout 300 python bench.py --precond jacobi --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/f_bench_cfg2_jacobi.json 2> $OUT/f_bench_cfg2_jacobi.err; echo "jacobi rc=$?"
python tools/bench_summary.py $OUT/f_bench_default.json $OUT/f_bench_cfg3.json $OUT/f_bench_cfg4.json $OUT/f_bench_cfg2_f32.json $OUT/f_bench_cfg2_jacobi.json
fi
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
python bench.py $ARGS > $OUT/f_plain.json 2> $OUT/f_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/f_launches.csv python bench.py $ARGS > $OUT/f_ncu_launches.log 2>&1
echo "launch list rc=$?"
for K in k_cg_step k_cg_update k_zu_march; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 2 -f -o $OUT/f_$K python bench.py $ARGS > $OUT/f_ncu_$K.log 2>&1
  echo "$K rc=$?"; export_rep f_$K
done
ARGS3="--workload cfg3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
python bench.py $ARGS3 > $OUT/f3_plain.json 2> $OUT/f3_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/f3_launches.csv python bench.py $ARGS3 > $OUT/f3_ncu_launches.log 2>&1
echo "cfg3 launch list rc=$?"
for K in k_cg_step k_zu_march; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 2 -f -o $OUT/f3_$K python bench.py $ARGS3 > $OUT/f3_ncu_$K.log 2>&1
  echo "cfg3 $K rc=$?"; export_rep f3_$K
done
du -sh $OUT
