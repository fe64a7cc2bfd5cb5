#!/bin/bash
# one gpurun call: parity tests, bench, host-side profile, ncu launch list + full capture
# usage: scripts/gpu_round.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -6 gpurun_out/${tag}_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
LINKB200_PROFILE_HOST=1 timeout 300 python bench.py --no-encoder --no-cpu-baseline --steps 50 > /dev/null 2> gpurun_out/${tag}_host_block.txt
LINKB200_PROFILE_HOST=1 timeout 300 python bench.py --workload encoder --no-cpu-baseline --steps 20 > /dev/null 2> gpurun_out/${tag}_host_enc.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-encoder --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"link_|conv_tc|linear_ln|kmap_query|plan_|table_insert" -s 44 -c 16 -f -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-encoder --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1; tail -3 gpurun_out/${tag}_ncu.log
ls -la gpurun_out/
