#!/bin/bash
# first GPU call for the r2-prep branch: the pieces written without a GPU
#   * tensor.UploadRing (+ bench.py --e2e-ring): does the upload of scan i+1 overlap scan i?
#   * csrc/iou3d.cu lk_boxes_iou_bev: device instantiation of the host-pinned pair arithmetic
tag=${1:-r2prep}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "upload_ring or rotated_iou or from_host" > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
timeout 200 python bench.py --no-cpu-baseline --no-encoder --e2e-ring > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.load(open('gpurun_out/${tag}_bench.json'))
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'e2e ring', d.get('e2e_ring'))
PY
timeout 120 python - <<'PY'
# rotated IoU: fused kernel vs the tensor-op formulation, n = 1000 boxes
import math, time, torch
from link_b200 import iou3d
g = torch.Generator().manual_seed(0)
b = torch.cat([(torch.rand(1000, 2, generator=g) - 0.5) * 60, torch.zeros(1000, 1), torch.rand(1000, 3, generator=g) * 4 + 0.5,
               (torch.rand(1000, 1, generator=g) - 0.5) * 6], 1).cuda()
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
k = iou3d.boxes_iou_bev(b, b)
t = torch.cat([iou3d._overlap_block(b[r:r + 256], b) for r in range(0, 1000, 256)])
area = (b[:, 3] * b[:, 4])
t = t / (area[:, None] + area[None, :] - t).clamp(min=1e-8)
print('kernel vs tensor ops max diff', float((k - t).abs().max()))
print('kernel ms', timed(lambda: iou3d.boxes_iou_bev(b, b)), 'rotate_nms ms', timed(lambda: iou3d.rotate_nms(b, torch.rand(1000, device='cuda'), 0.2)))
PY
