#!/bin/bash
# pre-aggregation kernel A/B: the kernel alone, default library + variant libraries, one process
tag=${1:-x}
mkdir -p gpurun_out
out=gpurun_out/${tag}_preagg_ab.jsonl; : > $out
timeout 150 python scripts/preagg_ab.py --tag ring --libs link_b200/liblinkb200_*.so >> $out 2>gpurun_out/${tag}_err.log
LINKB200_PREAGG=smem timeout 100 python scripts/preagg_ab.py --tag smem >> $out 2>>gpurun_out/${tag}_err.log
cut -c1-200 $out; tail -5 gpurun_out/${tag}_err.log
