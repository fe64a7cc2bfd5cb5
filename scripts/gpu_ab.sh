#!/bin/bash
# A/B: main library vs a variant library, tests + kernel microbench + bench
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -4 gpurun_out/${tag}_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('MAIN ms/step',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],'enc ms',d['encoder']['ms_per_step'])
for k,v in d['kernels'].items(): print(f"  {k:24s} {v['avg_us']:7.1f} us")
PY
for lib in link_b200/liblinkb200_*.so; do
  LINKB200_LIB=$PWD/$lib timeout 600 python -m pytest tests -m gpu -x -q -k "block or preagg or tselk or full_size" > gpurun_out/${tag}_tests_var.log 2>&1; tail -2 gpurun_out/${tag}_tests_var.log
  LINKB200_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu-baseline --no-encoder > gpurun_out/${tag}_bench_var.json 2>> gpurun_out/${tag}_bench.err; python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_var.json'))
print('VARIANT $lib ms/step',d['ms_per_step'])
for k,v in d['kernels'].items(): print(f"  {k:24s} {v['avg_us']:7.1f} us")
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|link_apply|linear_ln" -s 10 -c 6 -f -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-encoder --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1; tail -2 gpurun_out/${tag}_ncu.log
