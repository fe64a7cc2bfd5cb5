#!/bin/bash
# BASELINE config 5 at N GPUs (replicas, max-over-ranks time): bash scripts/sweep_multi.sh N [extra sweep args]
N=${1:-2}; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540 + N)) \
  scripts/sweep.py --skip-det --out gpurun_out/r02_sweep_${N}gpu.json "$@" 2>&1 | grep "^{" | tail -3 | cut -c1-220
