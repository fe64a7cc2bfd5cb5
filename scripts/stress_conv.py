"""Stress the tensor-core conv against the FFMA conv on many random maps (hunting rare races):
    python scripts/stress_conv.py [iterations]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from link_b200 import SparseTensor, _capi
import link_b200.nn.functional.conv as cv
from link_b200.utils.synthetic import random_voxels
import ctypes as C

dev = torch.device('cuda:0')
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(0)
L = _capi.lib()
worst = 0.0
bad = 0
for it in range(iters):
    n = int(rng.integers(200, 30000))
    c = int(rng.choice([32, 64, 128]))
    coords = torch.from_numpy(random_voxels(n, int(rng.integers(16, 48)), seed=it)).to(dev)
    n = coords.shape[0]
    st = SparseTensor(torch.randn(n, c, device=dev), coords, 1)
    km = cv.build_kernel_map(st, (3, 3, 3), (1, 1, 1), (1, 1, 1), want_plan=bool(it % 2))
    w = torch.nn.Parameter(torch.randn(27, c, c, device=dev) * 0.05)
    got = cv._conv_fwd(st.F, w, km.nbr, n, kmap=km if it % 2 else None)
    ref = torch.empty(n, c, device=dev)
    ep = _capi.ConvEpilogue()
    _capi.check(L.lk_conv_fwd_ex(_capi.ptr(st.F), _capi.ptr(w.detach()), _capi.ptr(km.nbr), n, 27, c, c, C.byref(ep),
                                 _capi.ptr(ref), _capi.stream()), 'ffma')
    err = float((got - ref).abs().max())
    scale = float(ref.abs().max())
    worst = max(worst, err / max(scale, 1e-6))
    if err > 2e-5 * max(scale, 1.0):
        bad += 1
        print('MISMATCH it', it, 'n', n, 'c', c, 'err', err, 'scale', scale, flush=True)
print('iterations', iters, 'mismatches', bad, 'worst relative error', worst)
