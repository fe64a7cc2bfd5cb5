"""Small instances of every pipeline with hand-rolled synchronisation, for compute-sanitizer
(scripts/sanitize.sh): the tcgen05 conv (mbarrier ring + TMEM A stages + bulk-copied weights), the
tcgen05 weight gradient (MN-major staging, two barrier families), the tcgen05 Linear+LayerNorm,
the ring pre-aggregation kernel, the two-stream native block executor (forward and backward) and the
native encoder executor."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from link_b200 import SparseTensor
from link_b200.elk import ELKBlock
from link_b200.nn.functional import conv as conv_mod
from link_b200.utils.synthetic import random_voxels

dev = torch.device('cuda:0')
torch.manual_seed(0)
n = int(os.environ.get('SANITIZE_N', '1500'))
coords = torch.from_numpy(random_voxels(n, 20, seed=1, batch=2)).to(dev)
for c, prec in [(32, 'fp32'), (64, 'fp32'), (64, 'tf32'), (128, 'fp32')]:
    conv_mod.set_precision(prec)
    x = SparseTensor(torch.randn(len(coords), c, device=dev, requires_grad=True), coords, 1)
    x.cmaps[x.stride] = x.coords
    w = torch.randn(27, c, c, device=dev, requires_grad=True)
    import link_b200.nn.functional as F
    y = F.conv3d(x, w, 3)
    y.F.square().sum().backward()           # forward, dgrad (same kernel) and the tcgen05 wgrad
    torch.cuda.synchronize()
    print(f'conv C={c} {prec}: ok', float(w.grad.abs().sum()), flush=True)
conv_mod.set_precision('fp32')
for c, groups, s, r in [(64, 2, 7, 3), (32, 1, 3, 2)]:
    blk = ELKBlock(c, c, groups=groups, baseop='cos').to(dev)
    f = torch.randn(len(coords), c, device=dev)
    with torch.no_grad():
        out = blk.eval()(SparseTensor(f.clone(), coords, 1), s, r).F       # native executor, two streams
    fr = f.clone().requires_grad_(True)
    blk.train()(SparseTensor(fr, coords, 1), s, r).F.sum().backward()     # fused backward kernels
    torch.cuda.synchronize()
    print(f'block C={c} ({r}x{s})^3: ok', float(out.abs().sum()), float(fr.grad.abs().sum()), flush=True)
# fused training-mode BatchNorm + shortcut + ReLU (double accumulators, odd channel-group counts), forward and backward
import link_b200.nn.functional as F2
for nn_, cc in [(4097, 20), (3000, 128), (257, 1024), (2, 4)]:
    bn = torch.nn.BatchNorm1d(cc).to(dev).train()
    xb = torch.randn(nn_, cc, device=dev, requires_grad=True)
    rb = torch.randn(nn_, cc, device=dev, requires_grad=True)
    F2.batch_norm_act(xb, bn, True, rb).square().sum().backward()
    torch.cuda.synchronize()
    print(f'bn n={nn_} c={cc}: ok', float(xb.grad.abs().sum()), flush=True)
# the native encoder executor: five streams, four sort chains, one workspace arena carved for n0 rows per level
from link_b200.linkencoder import ELKEncoder
enc = ELKEncoder(num_classes=19, cr=0.5, baseop='cos', r=3, s=7, groups=2).to(dev).eval()
xs = torch.randn(len(coords), 4, device=dev)
with torch.no_grad():
    for _ in range(2):
        logits = enc(SparseTensor(xs, coords, 1))
torch.cuda.synchronize()
assert '_lk_enc_native' in enc.__dict__
print('native encoder: ok', float(logits.abs().sum()), flush=True)
print('sanitize cases done')
