"""Micro-benchmark of the pre-aggregation PATH (everything between F_input and the block
output) on the bench scan: every kernel alone and the chains, each captured into a CUDA graph over
rotating buffers (> L2) and replayed; CUDA events around the replay.
usage: python scripts/bench_path.py [--voxels 120000] [--c 64] [--reps 4] [--json out.json]"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def graph_time(fn_list, n_launch, reps=3):
    """fn_list: callables taking (i, stream).  Captures n_launch rounds, replays, returns us per round."""
    from link_b200 import _capi
    for i in range(2):
        for fn in fn_list:
            fn(i, _capi.stream())
    torch.cuda.synchronize()
    gr, cap = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.graph(gr, stream=cap):
        cs = _capi.stream()
        for i in range(n_launch):
            for fn in fn_list:
                fn(i, cs)
    gr.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, 1e3 * e0.elapsed_time(e1) / n_launch)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--voxels', type=int, default=120_000)
    ap.add_argument('--c', type=int, default=64)
    ap.add_argument('--groups', type=int, default=2)
    ap.add_argument('--s', type=int, default=7)
    ap.add_argument('--r', type=int, default=3)
    ap.add_argument('--nbuf', type=int, default=6)
    ap.add_argument('--rounds', type=int, default=24)
    ap.add_argument('--json', default=None)
    a = ap.parse_args()
    from link_b200 import SparseTensor, _capi
    from link_b200.elk import block_index, _kernel_gen
    from link_b200.utils.synthetic import kitti_like_voxels
    dev = torch.device('cuda:0')
    c3, _ = kitti_like_voxels(a.voxels, seed=0)
    coords = torch.from_numpy(np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32)).to(dev)
    n, c, k = coords.shape[0], a.c, 2
    st = SparseTensor(torch.zeros(n, 1, device=dev), coords, 1)
    bi = block_index(st, a.s)
    m = bi.m
    nbr = bi.neighbors(a.r)
    r3 = nbr.shape[1]
    torch.manual_seed(0)
    w = (torch.randn(c // a.groups, 3, device=dev) * 0.3).contiguous()
    gen = _kernel_gen('cos', c, w, None, 1.0)
    nb = a.nbuf
    fin = [torch.randn(n, c, device=dev) for _ in range(nb)]
    local = [torch.randn(n, c, device=dev) for _ in range(nb)]
    out = [torch.empty(n, c, device=dev) for _ in range(nb)]
    sums = torch.zeros(n, k * c, device=dev)
    mean = torch.zeros(n, k * c, device=dev)
    g1, b1, g2, b2 = (torch.rand(c, device=dev) + 0.5 for _ in range(4))
    L = _capi.lib()
    P = _capi.ptr

    def zero(i, s):
        _capi.check(L.lk_zero_rows(P(sums), P(bi.num), n, k * c, s), 'zero')

    def preagg(i, s):
        _capi.check(L.lk_link_preagg_seg_fwd(P(fin[i % nb]), P(coords), P(bi.order), P(bi.sorted_rank), n, C.byref(gen),
                                             P(sums), s), 'preagg')

    def wmean(i, s):
        _capi.check(L.lk_link_window_mean_seg(P(sums), P(bi.seg), P(nbr), P(bi.num), n, r3, k * c, P(mean), s), 'wm')

    def apply_(norm):
        def f(i, s):
            _capi.check(L.lk_link_apply_fwd(P(mean), P(fin[i % nb]), P(coords), P(bi.idx_query), n, C.byref(gen), norm,
                                            P(local[i % nb]) if norm else None, P(g1), P(b1), P(g2), P(b2), P(out[i % nb]), s), 'apply')
        return f

    # SURVEY 8d path bytes (pre-LayerNorm output) and the in-block variant (+ local row read)
    path_bytes = n * (2 * 4 * c + 2 * 16 + 2 * 4) + m * (2 * (4 * k * c + 4))
    block_bytes = path_bytes + n * 4 * c
    res = {'n': n, 'm': m, 'c': c, 'path_bytes': path_bytes, 'block_bytes': block_bytes}
    cases = {
        'preagg': [preagg],
        'window_mean': [wmean],
        'apply_plain': [apply_(0)],
        'apply_norm': [apply_(1)],
        'zero': [zero],
        'path (zero+preagg+wmean+apply_plain)': [zero, preagg, wmean, apply_(0)],
        'path_nozero (preagg+wmean+apply_plain)': [preagg, wmean, apply_(0)],
        'block_ (zero+preagg+wmean+apply_norm)': [zero, preagg, wmean, apply_(1)],
    }
    only = os.environ.get('BENCH_PATH_ONLY')
    for name, fns in cases.items():
        if only and only not in name:
            continue
        us = graph_time(fns, a.rounds)
        res[name] = us
        extra = ''
        if name.startswith('path'):
            extra = f'  -> {path_bytes / us / 1e3:.0f} GB/s on {path_bytes / 1e6:.1f} MB'
        if name.startswith('block_'):
            extra = f'  -> {block_bytes / us / 1e3:.0f} GB/s on {block_bytes / 1e6:.1f} MB'
        print(f'{name:46s} {us:8.2f} us{extra}', flush=True)
    if a.json:
        json.dump(res, open(a.json, 'w'), indent=1)


if __name__ == '__main__':
    main()
