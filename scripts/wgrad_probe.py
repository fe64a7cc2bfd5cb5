"""lk_conv_wgrad_tc against a float64 contraction, and its time vs the FFMA kernel (lk_conv_bwd_weight)
for 1 / 2 / 4 / all accumulator slots per CTA.  (The first version of this script also probed the
shared-memory layouts tcgen05 accepts for MN-major tf32 operands: profiles/r02_wgrad_probe.txt.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from link_b200 import SparseTensor, _capi
from link_b200.nn.functional.conv import build_kernel_map
from link_b200.utils.synthetic import kitti_like_voxels

dev = torch.device('cuda:0')


def kmap(n, cin):
    c3, _ = kitti_like_voxels(n, seed=3)
    coords = np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32)
    st = SparseTensor(torch.zeros(len(coords), cin, device=dev), torch.from_numpy(coords).to(dev), 1)
    return build_kernel_map(st, (3, 3, 3), (1, 1, 1), (1, 1, 1), want_plan=True)


def run(km, x, g, cin, cout, slots=0):
    nbrp, perm, masks = km.wgrad_relation(False)
    K, rows = km.nbr.shape
    gw = torch.empty(K, cin, cout, device=dev)
    _capi.check(_capi.lib().lk_conv_wgrad_tc(_capi.ptr(x), _capi.ptr(g), _capi.ptr(nbrp), _capi.ptr(perm),
                                             _capi.ptr(masks), rows, K, cin, cout, _capi.ptr(gw), slots,
                                             _capi.stream()), 'wgrad')
    return gw


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for cin, cout in [(64, 64), (32, 32), (128, 128)]:
    km = kmap(8_000, cin)
    K, rows = km.nbr.shape
    x = torch.randn(km.n_in, cin, device=dev)
    g = torch.randn(rows, cout, device=dev)
    rel = km.nbr.cpu().numpy()
    want = np.zeros((K, cin, cout))
    xn, gn = x.double().cpu().numpy(), g.double().cpu().numpy()
    for k in range(K):
        hit = rel[k] >= 0
        want[k] = xn[rel[k][hit]].T @ gn[hit]
    got = run(km, x, g, cin, cout).cpu().numpy()
    print(f'C {cin}x{cout}: max err / max|want| = {np.abs(got - want).max() / np.abs(want).max()}', flush=True)
for n, cin, cout in [(120_000, 64, 64), (160_000, 64, 64), (40_000, 128, 128), (160_000, 32, 32), (15_000, 64, 64)]:
    km = kmap(n, cin)
    K, rows = km.nbr.shape
    x = torch.randn(km.n_in, cin, device=dev)
    g = torch.randn(rows, cout, device=dev)
    gw = torch.empty(K, cin, cout, device=dev)
    t_old = timeit(lambda: _capi.check(_capi.lib().lk_conv_bwd_weight(_capi.ptr(x), _capi.ptr(g), _capi.ptr(km.nbr), rows, K,
                                                                      cin, cout, _capi.ptr(gw), _capi.stream()), 'old'))
    km.wgrad_relation(False)
    line = f'N {rows} C {cin}x{cout}: ffma {t_old:.1f} us'
    for slots in (0, 1, 2, 4):
        line += f' | tc slots={slots}: {timeit(lambda: run(km, x, g, cin, cout, slots)):.1f} us'
    km._wgrad.clear()
    line += f' | prepass {timeit(lambda: (km._wgrad.clear(), km.wgrad_relation(False))):.1f} us'
    print(line, flush=True)
