"""Tensor-core sparse conv: 3xTF32 ('fp32') vs single-pass TF32 ('tf32') -- time per launch (CUDA-graph
replay of 8 launches over rotating inputs) and max error against a float64 contraction."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from link_b200 import SparseTensor, _capi
from link_b200.nn.functional import conv as conv_mod
from link_b200.utils.synthetic import kitti_like_voxels

dev = torch.device('cuda:0')


def graph_us(fn, n=8, reps=3):
    fn(0); torch.cuda.synchronize()
    gr, cap = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.graph(gr, stream=cap):
        for i in range(n):
            fn(i)
    gr.replay(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, 1e3 * e0.elapsed_time(e1) / n)
    return best


for n, c in [(120_000, 64), (160_000, 64), (120_000, 32), (40_000, 128), (15_000, 64)]:
    c3, _ = kitti_like_voxels(n, seed=3)
    coords = np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32)
    st = SparseTensor(torch.zeros(len(coords), c, device=dev), torch.from_numpy(coords).to(dev), 1)
    km = conv_mod.build_kernel_map(st, (3, 3, 3), (1, 1, 1), (1, 1, 1), want_plan=True)
    K, rows = km.nbr.shape
    xs = [torch.randn(rows, c, device=dev) for _ in range(4)]
    w = torch.nn.Parameter(torch.randn(K, c, c, device=dev) / np.sqrt(c * 4))
    rel = km.nbr.cpu().numpy()
    want = np.zeros((rows, c))
    xn, wn = xs[0].double().cpu().numpy(), w.detach().double().cpu().numpy()
    for k in range(K):
        hit = rel[k] >= 0
        want[hit] += xn[rel[k][hit]] @ wn[k]
    line = f'N {rows} C {c}:'
    for prec in ('fp32', 'tf32'):
        conv_mod.set_precision(prec)
        with torch.no_grad():
            got = conv_mod._conv_fwd(xs[0], w, km.nbr, rows, kmap=km).double().cpu().numpy()
            us = graph_us(lambda i: conv_mod._conv_fwd(xs[i % 4], w, km.nbr, rows, kmap=km))
        err = np.abs(got - want).max() / np.abs(want).max()
        line += f'  {prec}: {us:7.1f} us  err/max {err:.2e}'
    conv_mod.set_precision('fp32')
    xb = [t.to(torch.bfloat16) for t in xs]
    if c <= 64:
        with torch.no_grad():
            got = conv_mod._conv_fwd(xb[0], w, km.nbr, rows, kmap=km).double().cpu().numpy()
            us = graph_us(lambda i: conv_mod._conv_fwd(xb[i % 4], w, km.nbr, rows, kmap=km))
        xq = xb[0].double().cpu().numpy()
        wantq = np.zeros((rows, c))
        for k in range(K):
            hit = rel[k] >= 0
            wantq[hit] += xq[rel[k][hit]] @ wn[k]
        line += f'  bf16 rows: {us:7.1f} us  err/max {np.abs(got - wantq).max() / np.abs(wantq).max():.2e}'
    print(line, flush=True)
