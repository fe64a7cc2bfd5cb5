"""How much of the event-timed block step is the DEVICE waiting for the host?  The same step (fresh
SparseTensor, every index map rebuilt) timed (a) as bench.py does: flush, event, python enqueue, event;
(b) with a long device-side delay in front of the first event, so that the step's launches are queued
before the device reaches them; (c) captured once into a CUDA graph and replayed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from link_b200 import SparseTensor
from link_b200.elk import ELKBlock
from link_b200.nn.functional import _index
from link_b200.utils.synthetic import kitti_like_voxels
dev = torch.device('cuda:0')
c3, _ = kitti_like_voxels(120_000, seed=0)
ch = np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32)
coords = torch.from_numpy(ch).to(dev)
lo, hi = ch.min(0), ch.max(0)
torch.manual_seed(0)
blk = ELKBlock(64, 64, groups=2, baseop='cos').to(dev).eval()
feats = torch.randn(len(ch), 64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
big = torch.empty(1536 << 20, dtype=torch.uint8, device=dev)

def step(f):
    st = SparseTensor(f, coords, 1)
    _index.set_coord_bounds(st.kmaps, lo, hi)
    with torch.no_grad():
        return blk(st, 7, 3).F

def timed(pre, n=30, warm=5):
    ts = []
    for k in range(warm + n):
        f = feats.clone()
        pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(f); e1.record()
        torch.cuda.synchronize()
        if k >= warm: ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts)), float(np.median(ts)), float(np.min(ts))
print('(a) 256 MB flush before the event   : mean %.4f median %.4f min %.4f ms' % timed(lambda: flush.zero_()))
print('(b) 1.5 GB memset before the event  : mean %.4f median %.4f min %.4f ms' % timed(lambda: big.zero_()))
# (c) graph replay
f = feats.clone()
for _ in range(3): step(f.clone())
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
fin = feats.clone()
with torch.cuda.graph(g):
    out = step(fin)
ts = []
for k in range(35):
    fin.copy_(feats); flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    if k >= 5: ts.append(e0.elapsed_time(e1))
print('(c) CUDA-graph replay of the step   : mean %.4f median %.4f min %.4f ms' % (np.mean(ts), np.median(ts), np.min(ts)))
