"""Debug aid: training-mode ELKBlock gradients (fused backward on/off) against the oracle's CPU autograd.
usage: python scripts/debug_block_grads.py op C groups s r n extent"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import link_oracle as O


def main():
    import link_b200.elk as elk
    from link_b200 import SparseTensor
    from link_b200.elk import ELKBlock
    from link_b200.utils.synthetic import random_voxels
    a = sys.argv[1:]
    op, C, groups, s, r, n, extent = a[0], int(a[1]), int(a[2]), int(a[3]), int(a[4]), int(a[5]), int(a[6])
    dev = torch.device('cuda:0')
    coords = random_voxels(n, extent, seed=C + s + r, batch=2)
    n = len(coords)
    torch.manual_seed(C + s)
    blk = ELKBlock(C, C, groups=groups, baseop=op)
    feats = torch.randn(n, C)
    go = torch.randn(n, C, generator=torch.Generator().manual_seed(2))
    p = {k: v.detach().clone().requires_grad_(True) for k, v in blk.state_dict().items()}
    f_cpu = feats.clone().requires_grad_(True)
    o, parts = O.elk_block_forward(f_cpu, coords, 1, p, s, r, op, groups, return_parts=True)
    for t in (parts['F_input'], parts['local']):
        t.retain_grad()
    o.backward(go)
    blk = blk.to(dev).train()
    for fused in (True, False):
        elk.FUSED_BACKWARD = fused
        blk.zero_grad(set_to_none=True)
        f = feats.to(dev).requires_grad_(True)
        out = blk(SparseTensor(f, torch.from_numpy(coords).to(dev), 1), s, r).F
        out.backward(go.to(dev))
        print(f'fused={fused}: out {float((out.detach().cpu() - o.detach()).abs().max()):.2e}  '
              f'd feats {float((f.grad.cpu() - f_cpu.grad).abs().max()):.2e} (scale {float(f_cpu.grad.abs().max()):.2e})')
        for k, v in blk.named_parameters():
            ref = p[k].grad
            print(f'    {k:22s} {float((v.grad.cpu() - ref).abs().max()):.2e} (scale {float(ref.abs().max()):.2e})')


if __name__ == '__main__':
    main()
