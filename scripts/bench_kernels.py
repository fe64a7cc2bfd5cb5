"""Per-kernel microbenchmark on the BASELINE-size scan (CUDA events, L2 flushed between reps).
    python scripts/bench_kernels.py [--voxels 120000] [--c 64] [--reps 20]
Prints one JSON line per kernel: avg/min microseconds and algorithmic GB/s."""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from link_b200 import SparseTensor, _capi
from link_b200.elk import ELKBlock, block_index, link_aggregate, _pre_mix_fused
from link_b200.nn.functional import _index
import link_b200.nn.functional as F
from link_b200.utils.synthetic import kitti_like_voxels


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--voxels', type=int, default=120_000)
    ap.add_argument('--c', type=int, default=64)
    ap.add_argument('--reps', type=int, default=20)
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    c3, _ = kitti_like_voxels(a.voxels, seed=0)
    coords = torch.from_numpy(np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1)).to(dev)
    n, c = coords.shape[0], a.c
    torch.manual_seed(0)
    blk = ELKBlock(c, c, groups=2, baseop='cos').to(dev).eval()
    x = torch.randn(n, c, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = SparseTensor(x, coords, 1)
    bi = block_index(st, 7)
    m = bi.m
    with torch.no_grad():
        local = blk.local_mix(st).F
        f_in = _pre_mix_fused(blk.pre_mix, x)
    kmap = st.kmaps[((1, 1, 1), (3, 3, 3), (1, 1, 1), (1, 1, 1))]

    _capi.TIMERS = None
    w = blk.pos_weight[0].weight
    nrm = (blk.norm.weight, blk.norm.bias, blk.norm_local.weight, blk.norm_local.bias)
    # individual library calls through the same python wrappers the block uses
    _capi.TIMERS = {}
    with torch.no_grad():
        for r in range(a.reps + 3):
            flush.zero_()
            link_aggregate(f_in, coords, bi, 3, 'cos', w, None, 1.0, local, nrm)
            _pre_mix_fused(blk.pre_mix, x)
            F.conv_bn_act(st, blk.local_mix[0])
            st2 = SparseTensor(x, coords, 1)
            _index.set_coord_bounds(st2.kmaps, coords.min(0).values.tolist(), coords.max(0).values.tolist())
            block_index(st2, 7).neighbors(3)
            F.build_kernel_map(st2, (3, 3, 3), (1, 1, 1), (1, 1, 1))
    torch.cuda.synchronize()
    for name, lst in _capi.TIMERS.items():
        ts = [e0.elapsed_time(e1) * 1e3 for e0, e1, _ in lst][3:]
        nb = lst[-1][2]
        print(json.dumps({name: {'avg_us': float(np.mean(ts)), 'min_us': float(np.min(ts)),
                                 'gbs_at_min': nb / (np.min(ts) * 1e-6) / 1e9 if nb else None}}), flush=True)
    _capi.TIMERS = None
    print(json.dumps({'n': n, 'm': m, 'c': c}))


if __name__ == '__main__':
    main()
