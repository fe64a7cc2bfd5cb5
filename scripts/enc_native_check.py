"""Native encoder executor (lk_elk_encoder_fwd) vs the per-layer python path: same logits (bit-exact
expected: same kernels, same order of the per-row sums), and wall time per scan of both."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import link_b200.linkencoder as le
from link_b200 import SparseTensor
from link_b200.nn.functional import _index

dev = torch.device('cuda:0')
for nvox in (120_000, 20_000):
    c, f = bench.make_scan(nvox, seed=1)
    coords, feats = torch.from_numpy(c).to(dev), torch.from_numpy(f).to(dev)
    lo, hi = c.min(0), c.max(0)
    torch.manual_seed(0)
    net = le.ELKEncoder(num_classes=19, cr=1.0, baseop='cos', r=3, s=7, groups=2).to(dev).eval()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.uniform_(-0.2, 0.2); m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.2, 0.2)

    def fwd():
        st = SparseTensor(feats.clone(), coords, 1)
        _index.set_coord_bounds(st.kmaps, lo, hi)
        with torch.no_grad():
            return net(st)

    res = {}
    for name, nat, ov in (('python', False, True), ('native', True, True), ('native, no branch overlap', True, False)):
        le.NATIVE_ENCODER, le.BRANCH_OVERLAP = nat, ov
        net.__dict__.pop('_lk_enc_native', None)
        out = fwd()
        for _ in range(4):
            fwd()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            fwd()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        res[name] = out
        print(f'N={len(c)} {name:28s} wall {(time.perf_counter() - t0) * 50:.3f} ms/scan  (host enqueue {t_host * 50:.3f})', flush=True)
    d = (res['native'] - res['python']).abs().max().item()
    d2 = (res['native, no branch overlap'] - res['python']).abs().max().item()
    print(f'N={len(c)} max|native - python| = {d:.3e} / {d2:.3e}   |logits| max {res["python"].abs().max().item():.3f}', flush=True)
