"""One forward (and optionally one forward+backward) of the spconv-free detection backbone on a
nuScenes-shaped synthetic grid, for ncu launch lists:  python scripts/det_fwd.py [--bwd]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from link_b200.scn import SpMiddleResNetFHDELKv3
from link_b200.utils.synthetic import lidar_scan


def grid():
    pts = np.concatenate([lidar_scan(seed=10 + k, beams=32, azimuths=1100) for k in range(10)], 0)
    rng = [-54, -54, -5, 54, 54, 3]
    vs = np.array([0.075, 0.075, 0.2])
    keep = np.all((pts[:, :3] >= rng[:3]) & (pts[:, :3] < rng[3:]), axis=1)
    ijk = np.unique(np.floor((pts[keep, :3] - np.array(rng[:3])) / vs).astype(np.int32), axis=0)
    ijk = ijk[np.random.default_rng(0).permutation(len(ijk))[:120_000]]
    return np.concatenate([np.zeros((len(ijk), 1), np.int32), ijk[:, ::-1]], 1).astype(np.int32)


def main():
    dev = torch.device('cuda:0')
    idx = torch.from_numpy(grid()).to(dev)
    feats = torch.randn(idx.shape[0], 5, device=dev)
    torch.manual_seed(0)
    net = SpMiddleResNetFHDELKv3(num_input_features=5, ds_factor=8).to(dev).eval()
    reps = 3
    if '--bwd' in sys.argv:
        net.train()
        for _ in range(reps):
            f = feats.clone().requires_grad_(True)
            net(f, idx, 1, [1440, 1440, 40])[0].square().mean().backward()
    else:
        with torch.no_grad():
            for _ in range(reps):
                net(feats, idx, 1, [1440, 1440, 40])
    torch.cuda.synchronize()
    print('ok', idx.shape[0])


if __name__ == '__main__':
    main()
