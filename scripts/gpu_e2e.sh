#!/bin/bash
# upload-overlap check: from_host tests + the default bench line
tag=${1:-x}
mkdir -p gpurun_out
timeout 80 python -m pytest tests -m gpu -x -q -k "from_host" > gpurun_out/${tag}_tests.log 2>&1; tail -2 gpurun_out/${tag}_tests.log
timeout 110 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'lat',d['e2e']['latency_ms_per_step'],'enc',d.get('encoder',{}).get('ms_per_step'),d.get('encoder',{}).get('e2e'),'roof',d['roofline']['avg_us'],d['roofline']['frac'])
PY
