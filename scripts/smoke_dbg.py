"""Debug aid: ELKBlock forward vs the CPU oracle, optionally forcing the native executor at C=16."""
import sys, os
sys.path.insert(0, '.')
import numpy as np, torch
from link_b200 import SparseTensor, elk
from link_b200.elk import ELKBlock
from link_b200.utils.synthetic import random_voxels
from oracle import link_oracle as O
dev = torch.device('cuda:0')
force = '--force-native' in sys.argv
for baseop, groups, C, s, r in [('cos_x', 1, 16, 3, 2), ('cos', 2, 16, 5, 3)]:
    coords = random_voxels(8000, 64, seed=0)
    torch.manual_seed(0)
    blk = ELKBlock(C, C, groups=groups, baseop=baseop).eval()
    feats = torch.randn(len(coords), C)
    want = O.elk_block_forward(feats, coords, 1, {k: v.detach() for k, v in blk.state_dict().items()}, s, r, baseop, groups)
    blk = blk.to(dev)
    for rep in range(3):
        st = SparseTensor(feats.to(dev), torch.from_numpy(coords).to(dev), 1)
        with torch.no_grad():
            if force:
                scale = float(st.s[0]) if baseop == 'cos_x' else 1.0
                got = elk._forward_native(st, s, r, op=baseop, pre_mix=blk.pre_mix, conv=blk.local_mix[0],
                                          pos_weight=blk.pos_weight[0].weight, alpha=getattr(blk, 'alpha', None),
                                          coord_scale=scale, norm=blk.norm, norm_local=blk.norm_local)
            else:
                got = blk(st, s, r).F
        torch.cuda.synchronize()
        d = (got.cpu() - want).abs()
        print(baseop, C, 'rep', rep, 'max err', float(d.max()), 'n>4e-5', int((d > 4e-5).sum()))
