#!/bin/bash
# short block bench for the main library and every variant library: scripts/gpu_libs.sh <tag>
tag=$1
mkdir -p gpurun_out
for lib in link_b200/liblinkb200.so link_b200/liblinkb200_*.so; do
  [ -f "$lib" ] || continue
  name=$(basename $lib .so)
  LINKB200_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu-baseline --no-encoder --steps 30 > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err
  python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_$name.json'))
print('$name ms/step %.4f' % d['ms_per_step'], ' conv %.1f us' % d['kernels']['lk_conv_fwd']['avg_us'], ' plan %.1f' % d['kernels']['lk_conv_plan']['avg_us'])
PY
done
