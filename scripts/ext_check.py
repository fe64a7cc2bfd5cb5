import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from link_b200 import SparseTensor, _ext
import link_b200.nn.functional as F
from link_b200.utils.synthetic import random_voxels
dev = torch.device('cuda:0')
print('ext', _ext.module())
coords = torch.from_numpy(random_voxels(5000, 30, seed=1, batch=2)).to(dev)
res = {}
for use in (True, False):
    _ext.ENABLED = use; _ext._mod = None; _ext._tried = False
    torch.manual_seed(0)
    x = SparseTensor(torch.randn(len(coords), 64, device=dev, requires_grad=True), coords, 1)
    x.cmaps[x.stride] = x.coords
    w = torch.randn(27, 64, 32, device=dev, requires_grad=True)
    w2 = torch.randn(8, 32, 32, device=dev, requires_grad=True)
    y = F.conv3d(x, w, 3)
    z = F.conv3d(y, w2, 2, stride=2)
    bn = torch.nn.BatchNorm1d(32).to(dev).train()
    o = F.batch_norm_act(z.F, bn, True)
    o.square().sum().backward()
    res[use] = (o.detach(), x.F.grad.clone(), w.grad.clone(), w2.grad.clone(), bn.weight.grad.clone(), type(y.F.grad_fn).__name__, type(o.grad_fn).__name__)
print(res[True][5], res[True][6], '|', res[False][5], res[False][6])
for i, n in enumerate(['out', 'dx', 'dw', 'dw2', 'dgamma']):
    a, b = res[True][i], res[False][i]
    print(n, float((a - b).abs().max()), float(b.abs().max()))
