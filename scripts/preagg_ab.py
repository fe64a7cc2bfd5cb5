"""Pre-aggregation kernel alone: time (CUDA events around back-to-back launches over rotating
feature buffers larger than L2) and a check against a float64 numpy block sum.
    [LINKB200_LIB=...] [LINKB200_PREAGG=smem|ring] [LINKB200_PREAGG_Q=q] python scripts/preagg_ab.py [--voxels 120000 500000] [--c 64]"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from link_b200 import SparseTensor, _capi
from link_b200.elk import block_index, _kernel_gen
from link_b200.utils.synthetic import kitti_like_voxels


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--voxels', type=int, nargs='+', default=[120_000])
    ap.add_argument('--c', type=int, nargs='+', default=[64])
    ap.add_argument('--s', type=int, default=7)
    ap.add_argument('--reps', type=int, default=8)
    ap.add_argument('--tag', default='')
    ap.add_argument('--libs', nargs='*', default=[], help='extra variant libraries timed in the same process')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    L, stream = _capi.lib(), _capi.stream()
    libs = [('main', L)]
    for path in a.libs:
        V = C.CDLL(os.path.abspath(path))
        V.lk_link_preagg_seg_fwd.restype = L.lk_link_preagg_seg_fwd.restype
        V.lk_link_preagg_seg_fwd.argtypes = L.lk_link_preagg_seg_fwd.argtypes
        libs.append((os.path.basename(path), V))
    for nv in a.voxels:
        c3, _ = kitti_like_voxels(nv, seed=0)
        coords = torch.from_numpy(np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32)).to(dev)
        n = coords.shape[0]
        st = SparseTensor(torch.zeros(n, 1, device=dev), coords, 1)
        bi = block_index(st, a.s)
        m = bi.m
        for c in a.c:
            g = torch.Generator().manual_seed(c)
            w = (torch.randn(c // 2, 3, generator=g) * 0.05).to(dev)
            gen = _kernel_gen('cos', c, w, None, 1.0)
            nbuf = max(2, int(200e6 // (n * c * 4)) + 1)
            bufs = [torch.randn(n, c, device=dev) for _ in range(nbuf)]
            sums = torch.zeros(n, 2 * c, device=dev)

            for name, V in libs:
                def launch(i):
                    _capi.check(V.lk_link_preagg_seg_fwd(_capi.ptr(bufs[i % nbuf]), _capi.ptr(coords),
                                                         _capi.ptr(bi.order), _capi.ptr(bi.sorted_rank), n,
                                                         C.byref(gen), _capi.ptr(sums), stream), 'preagg')
                # correctness of one launch
                sums.zero_()
                launch(0)
                torch.cuda.synchronize()
                pos = (c3.astype(np.float32) @ w.cpu().numpy().T.astype(np.float32))
                pos = np.tile(pos, (1, 2)).astype(np.float64)
                fn = bufs[0].cpu().numpy().astype(np.float64)
                want = np.zeros((m, 2 * c))
                np.add.at(want, bi.idx_query.cpu().numpy(), np.concatenate([fn * np.cos(pos), fn * np.sin(pos)], 1))
                err = float(np.abs(sums[:m].cpu().numpy() - want).max())
                tail = float(sums[m:].abs().max()) if m < n else 0.0
                for i in range(nbuf):
                    launch(i)
                torch.cuda.synchronize()
                # the launches are captured into a CUDA graph: python + ctypes cost ~10 us per call,
                # which would otherwise bound the rate of a ~10 us kernel
                gr = torch.cuda.CUDAGraph()
                cap = torch.cuda.Stream()
                with torch.cuda.graph(gr, stream=cap):
                    cs = _capi.stream()
                    for i in range(nbuf * a.reps):
                        _capi.check(V.lk_link_preagg_seg_fwd(_capi.ptr(bufs[i % nbuf]), _capi.ptr(coords),
                                                             _capi.ptr(bi.order), _capi.ptr(bi.sorted_rank), n,
                                                             C.byref(gen), _capi.ptr(sums), cs), 'preagg')
                gr.replay()
                torch.cuda.synchronize()
                best = 1e9
                for _ in range(5):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    gr.replay()
                    e1.record()
                    torch.cuda.synchronize()
                    best = min(best, 1e3 * e0.elapsed_time(e1) / (nbuf * a.reps))
                nbytes = n * (4 * c + 16 + 4) + m * 4 * 2 * c
                print(json.dumps({'tag': a.tag, 'preagg': os.environ.get('LINKB200_PREAGG', 'ring'),
                                  'q': os.environ.get('LINKB200_PREAGG_Q'), 'lib': name,
                                  'n': n, 'm': m, 'c': c, 'us': round(best, 2), 'gbs': round(nbytes / best / 1e3, 1),
                                  'frac': round(nbytes / best / 1e3 / 6544.3, 3), 'max_abs_err': err, 'tail': tail}), flush=True)
            del bufs, sums


if __name__ == '__main__':
    main()
