"""Workload for `ncu --set full`: two fused LinK block forwards on the bench scan (N ~ 119k, C = 64,
cos (3x7)^3) and one 27-offset conv forward + backward (dgrad on the same tensor-core kernel, wgrad on
the tcgen05 weight-gradient kernel) on the same scan."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from link_b200 import SparseTensor
from link_b200.elk import ELKBlock
import link_b200.nn.functional as F
from link_b200.utils.synthetic import kitti_like_voxels

dev = torch.device('cuda:0')
c3, _ = kitti_like_voxels(120_000, seed=0)
coords = torch.from_numpy(np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32)).to(dev)
torch.manual_seed(0)
blk = ELKBlock(64, 64, groups=2, baseop='cos').to(dev).eval()
feats = torch.randn(coords.shape[0], 64, device=dev)
with torch.no_grad():
    for _ in range(2):
        blk(SparseTensor(feats.clone(), coords, 1), 7, 3)
torch.cuda.synchronize()
x = SparseTensor(feats.clone().requires_grad_(True), coords, 1)
x.cmaps[x.stride] = x.coords
w = torch.randn(27, 64, 64, device=dev, requires_grad=True)
for _ in range(2):
    y = F.conv3d(x, w, 3)
    y.F.square().sum().backward()
torch.cuda.synchronize()
print('profile_step done')
