#!/bin/bash
# Second evidence pass of round 1 on 1 GPU (after the shuffle-based 2-D CG kernels and the pinned e2e path):
# full GPU test suite, default bench line, ncu launch list and --set full captures of the new kernels.
OUT=gpurun_out
export_rep() {  # <name>
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
  ncu -i $OUT/$1.ncu-rep --page details > $OUT/$1_details.txt 2>/dev/null
  ncu -i $OUT/$1.ncu-rep --page source --csv > $OUT/$1_source.csv 2>/dev/null
  rm -f $OUT/$1.ncu-rep
}
timeout 420 python -m pytest tests -m gpu -x -q --durations=12 > $OUT/i_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/i_pytest.log
timeout 200 python bench.py > $OUT/i_bench_default.json 2> $OUT/i_bench_default.err; echo "default rc=$?"
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
python bench.py $ARGS > $OUT/i_plain.json 2> $OUT/i_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/i_launches.csv python bench.py $ARGS > $OUT/i_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_cg_step2d -s 4 -c 2 -f -o $OUT/i_k_cg_step2d python bench.py $ARGS > $OUT/i_ncu_k_cg_step2d.log 2>&1
echo "k_cg_step2d rc=$?"; export_rep i_k_cg_step2d
python -c "
import json; d=json.load(open('$OUT/i_bench_default.json'))
print('value %.4g e2e %.4g ms/pass %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
print({k: (round(v['avg_ms'], 4), round(v['frac'], 3)) for k, v in d['stages'].items()})"
du -sh $OUT
