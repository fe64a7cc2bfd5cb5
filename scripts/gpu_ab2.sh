#!/bin/bash
# tests + bench under two settings of an environment knob:  scripts/gpu_ab2.sh <tag> <VAR> <a> <b>
tag=$1; var=$2; a=$3; b=$4
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -4 gpurun_out/${tag}_tests.log
for v in $a $b; do
  env $var=$v timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_$v.json 2> gpurun_out/${tag}_bench_$v.err
  tail -2 gpurun_out/${tag}_bench_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_$v.json'))
print('$var=$v ms/step',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],'enc ms',d['encoder']['ms_per_step'], 'roof', d['roofline']['avg_us'], d['roofline']['frac'])
for k,x in d['kernels'].items(): print(f"  {k:24s} {x['avg_us']:7.1f} us")
PY
done
env $var=$b timeout 300 python -m pytest tests -m gpu -x -q -k "conv or encoder or unet or det" > gpurun_out/${tag}_tests_b.log 2>&1; tail -2 gpurun_out/${tag}_tests_b.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|link_apply|link_preagg|linear_ln" -s 10 -c 8 -f -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-encoder --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1; tail -2 gpurun_out/${tag}_ncu.log
