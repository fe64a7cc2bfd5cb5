#!/bin/bash
# one gpurun call: smoke, parity tests, default bench (with CPU + reference-GPU baselines), reference arm,
# ncu launch list + full capture, sweep
# usage: scripts/gpu_round.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -4 gpurun_out/${tag}_tests.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'enc ms',d['encoder']['ms_per_step'],'roof',d['roofline']['avg_us'],d['roofline']['frac'],'cpu',d.get('cpu_baseline',{}).get('value'),'refgpu',d.get('reference_gpu',{}).get('ms_per_step'))
for k,x in d['kernels'].items(): print(f"  {k:24s} {x['avg_us']:7.1f} us")
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err; cut -c1-200 gpurun_out/${tag}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-encoder --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"link_|conv_tc|linear_ln|kmap_query|plan_|radix|uniq|table_insert|block_neighbors" -s 60 -c 20 -f -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-encoder --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1; tail -2 gpurun_out/${tag}_ncu.log

