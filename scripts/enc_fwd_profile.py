"""Kernel table (torch.profiler) + host profile of the ELKEncoder inference forward (config 2) on the
bench scan: sum of kernel durations vs wall time per scan tells whether the step is device- or host-paced."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from link_b200 import SparseTensor
from link_b200.linkencoder import ELKEncoder
from link_b200.nn.functional import _index

dev = torch.device('cuda:0')
c, f = bench.make_scan(120_000, seed=0)
coords = torch.from_numpy(c).to(dev)
feats = torch.from_numpy(f).to(dev)
lo, hi = c.min(0), c.max(0)
torch.manual_seed(0)
net = ELKEncoder(num_classes=19, cr=1.0, baseop=bench.BASEOP, r=bench.R_BLK, s=bench.S_BLK, groups=bench.GROUPS).to(dev).eval()


def fwd():
    st = SparseTensor(feats.clone(), coords, 1)
    _index.set_coord_bounds(st.kmaps, lo, hi)
    with torch.no_grad():
        return net(st)


for _ in range(5):
    fwd()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    fwd()
torch.cuda.synchronize()
print('wall ms per scan', (time.perf_counter() - t0) * 100)
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        fwd()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=40, max_name_column_width=60))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    fwd()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(35)
