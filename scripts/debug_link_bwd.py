"""Debug aid: LinkAggregateFunction (fused forward + hand-written backward) against a float64 torch
autograd evaluation of the same function on the GPU, tensor by tensor.
usage: python scripts/debug_link_bwd.py [op C groups s r n extent]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def ref_aggregate(f, local, W, g1, b1, g2, b2, coords, idx, nbr, counts, op, groups):
    C = f.shape[1]
    pos = (coords[:, :3].double() @ W.t()).repeat(1, groups)
    cs, sn = torch.cos(pos), torch.sin(pos)
    planes = torch.cat([f * sn, f * cs], 1) if op == 'sin' else torch.cat([f * cs, f * sn], 1)
    M = counts.shape[0]
    S = torch.zeros(M, 2 * C, dtype=torch.float64, device=f.device).index_add(0, idx, planes)
    valid = (nbr >= 0)
    safe = nbr.clamp(min=0)
    tot = (counts[safe].double() * valid).sum(1)
    A = (S[safe] * valid[..., None]).sum(1) / tot[:, None]
    Ai = A[idx]
    y = Ai[:, :C] * cs - Ai[:, C:] * sn if op == 'sin' else Ai[:, :C] * cs + Ai[:, C:] * sn
    ln = torch.nn.functional.layer_norm
    return torch.relu(ln(y, (C,), g1, b1, 1e-6) + ln(local, (C,), g2, b2, 1e-6)), y


def main():
    from link_b200 import SparseTensor
    from link_b200 import elk
    from link_b200.utils.synthetic import random_voxels
    a = sys.argv[1:]
    op, C, groups, s, r, n, extent = (a[0], int(a[1]), int(a[2]), int(a[3]), int(a[4]), int(a[5]), int(a[6])) if len(a) >= 7 \
        else ('cos', 128, 4, 7, 3, 5000, 20)
    dev = torch.device('cuda:0')
    coords_h = random_voxels(n, extent, seed=C + s + r, batch=2)
    n = len(coords_h)
    coords = torch.from_numpy(coords_h).to(dev)
    st = SparseTensor(torch.zeros(n, C, device=dev), coords, 1)
    bi = elk.block_index(st, s)
    m = bi.m
    g = torch.Generator().manual_seed(0)
    mk = lambda *sh: torch.randn(*sh, generator=g).to(dev)
    f, local, W = mk(n, C), mk(n, C), mk(C // groups, 3) * 0.3
    g1, b1, g2, b2 = torch.rand(C, generator=g).to(dev) + 0.5, mk(C), torch.rand(C, generator=g).to(dev) + 0.5, mk(C)
    go = mk(n, C)
    leaves = [t.clone().requires_grad_(True) for t in (f, local, W, g1, b1, g2, b2)]
    out = elk.LinkAggregateFunction.apply(*leaves, coords, bi, r, op)
    out.backward(go)
    leaves64 = [t.double().clone().requires_grad_(True) for t in (f, local, W, g1, b1, g2, b2)]
    ref, y = ref_aggregate(*leaves64, coords, bi.idx_query.long(), bi.neighbors(r)[:m].long(), bi.counts[:m], op, groups)
    ref.backward(go.double())
    print(f'{op} C={C} groups={groups} s={s} r={r} n={n} m={m} max block {int(bi.counts[:m].max())}')
    print('forward   max|diff|', float((out.double() - ref).abs().max()))
    names = ['d f_input', 'd local', 'd pos_weight', 'd gamma1', 'd beta1', 'd gamma2', 'd beta2']
    for nm, a_, b_ in zip(names, leaves, leaves64):
        d = (a_.grad.double() - b_.grad).abs()
        sc = float(b_.grad.abs().max())
        print(f'{nm:14s} max|diff| {float(d.max()):.3e}  (scale {sc:.3e})  rows with err>1e-3*scale: '
              f'{int((d.reshape(d.shape[0], -1).max(1).values > 1e-3 * sc).sum())} / {d.shape[0]}')
        if nm in ('d f_input', 'd local') and float(d.max()) > 1e-3 * sc:
            bad = (d.max(1).values > 1e-3 * sc).nonzero().flatten()
            blk = bi.idx_query[bad].long()
            cnt = bi.counts[:m][blk]
            pos_in_sorted = torch.empty(n, dtype=torch.long, device=dev)
            pos_in_sorted[bi.order.long()] = torch.arange(n, device=dev)
            off = pos_in_sorted[bad] - bi.seg[:m][blk].long()
            print('   bad rows: block sizes', cnt[:12].tolist(), 'offset in block', off[:12].tolist(),
                  'bad channels of first row', (d[bad[0]] > 1e-3 * sc).nonzero().flatten()[:16].tolist())
            print('   offsets histogram (mod 32):', torch.bincount(off % 32, minlength=32).tolist())
            print('   min offset', int(off.min()), 'blocks<=32 affected', int((cnt <= 32).sum()))


if __name__ == '__main__':
    main()
