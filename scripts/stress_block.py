"""Stress the two-stream block executor against the single-stream order (hunting rare races):
    python scripts/stress_block.py [iterations]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from link_b200 import SparseTensor, elk
from link_b200.elk import ELKBlock
from link_b200.utils.synthetic import random_voxels

dev = torch.device('cuda:0')
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(1)
bad = 0
worst = 0.0
for it in range(iters):
    n = int(rng.integers(500, 40000))
    c = int(rng.choice([32, 64]))
    baseop = ['cos', 'sin', 'cos_x'][it % 3]
    groups = 1 if baseop == 'cos_x' else 2
    s, r = [(3, 2), (5, 3), (7, 3)][it % 3]
    coords_h = random_voxels(n, int(rng.integers(24, 64)), seed=it)
    torch.manual_seed(it)
    blk = ELKBlock(c, c, groups=groups, baseop=baseop).to(dev).eval()
    feats_h = torch.randn(len(coords_h), c).pin_memory()
    coords_hh = torch.from_numpy(coords_h).pin_memory()
    outs = []
    for single, host in ((True, False), (False, False), (False, True)):
        elk.SINGLE_STREAM = single
        with torch.no_grad():
            if host:
                st = SparseTensor.from_host(feats_h, coords_hh, 1, device=dev)
            else:
                st = SparseTensor(feats_h.to(dev), coords_hh.to(dev), 1)
            outs.append(blk(st, s, r).F.clone())
        # garbage-fill freed memory so that stale reads show up
        junk = torch.full((int(rng.integers(1, 8)) * 1_000_000,), float('nan'), device=dev)
        del junk
    torch.cuda.synchronize()
    for o in outs[1:]:
        err = float((o - outs[0]).abs().max()) if torch.isfinite(o).all() else float('inf')
        worst = max(worst, err)
        if not err < 2e-5:
            bad += 1
            print('MISMATCH it', it, 'n', len(coords_h), 'c', c, baseop, 'err', err, flush=True)
print('iterations', iters, 'mismatches', bad, 'worst abs diff', worst)
