"""Tensor-core vs FFMA sparse conv at small sizes (encoder levels 2-4), C = 64, K = 27, map prebuilt.
`tc` / `ffma`: 20 python launches back to back (host-paced at the small sizes); `tc graph`: the same
launch replayed from a CUDA graph (device time)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from link_b200 import SparseTensor
import link_b200.nn.functional.conv as cv
from link_b200.utils.synthetic import kitti_like_voxels

dev = torch.device('cuda:0')
c3, _ = kitti_like_voxels(120_000, seed=0)
for stride in (1, 2, 4, 8, 16):
    cc = np.unique(c3 // stride, axis=0).astype(np.int32)
    coords = torch.from_numpy(np.concatenate([cc, np.zeros((len(cc), 1), np.int32)], 1)).to(dev)
    n = coords.shape[0]
    st = SparseTensor(torch.randn(n, 64, device=dev), coords, 1)
    km = cv.build_kernel_map(st, (3, 3, 3), (1, 1, 1), (1, 1, 1), want_plan=True)
    w = torch.nn.Parameter(torch.randn(27, 64, 64, device=dev) * 0.05)
    res = {}
    for name, tc in (('tc', True), ('ffma', False)):
        cv.USE_TENSOR_CORES = tc
        for _ in range(3):
            cv._conv_fwd(st.F, w, km.nbr, n, kmap=km if tc else None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            cv._conv_fwd(st.F, w, km.nbr, n, kmap=km if tc else None)
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 20 * 1e3
    cv.USE_TENSOR_CORES = True
    gr, cap = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.graph(gr, stream=cap):
        for _ in range(8):
            cv._conv_fwd(st.F, w, km.nbr, n, kmap=km)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    res['graph'] = e0.elapsed_time(e1) / 8 * 1e3
    print(f'N={n:7d}  tc {res["tc"]:7.1f} us   tc graph {res["graph"]:7.1f} us   ffma {res["ffma"]:7.1f} us', flush=True)
