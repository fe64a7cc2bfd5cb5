"""Where does a training step go?  CUDA-event times plus a torch.profiler kernel table of
  * det : SpMiddleResNetFHDELKv3 forward / forward+backward on a nuScenes-shaped grid (config 4)
  * enc : ELKEncoder(cr=1.0, cos, (3x7)^3) forward+backward on 2 x 80k-voxel scans (config 3, one rank)
usage: python scripts/train_profile.py [det|enc] [--top 25]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def det_inputs(dev):
    from link_b200.utils.synthetic import lidar_scan
    pts = np.concatenate([lidar_scan(seed=10 + k, beams=32, azimuths=1100) for k in range(10)], 0)
    rng = [-54, -54, -5, 54, 54, 3]
    vs = np.array([0.075, 0.075, 0.2])
    keep = np.all((pts[:, :3] >= rng[:3]) & (pts[:, :3] < rng[3:]), axis=1)
    ijk = np.unique(np.floor((pts[keep, :3] - np.array(rng[:3])) / vs).astype(np.int32), axis=0)
    ijk = ijk[np.random.default_rng(0).permutation(len(ijk))[:120_000]]
    idx = np.concatenate([np.zeros((len(ijk), 1), np.int32), ijk[:, ::-1]], 1).astype(np.int32)
    return torch.randn(len(idx), 5, device=dev), torch.from_numpy(idx).to(dev)


def enc_inputs(dev, n_scans=2, n_vox=80_000):
    from link_b200.utils.synthetic import kitti_like_voxels
    cs, fs = [], []
    for b in range(n_scans):
        c3, f4 = kitti_like_voxels(n_vox, seed=20 + b)
        cs.append(np.concatenate([c3, np.full((len(c3), 1), b, np.int32)], 1))
        fs.append(f4)
    return torch.from_numpy(np.concatenate(fs).astype(np.float32)).to(dev), torch.from_numpy(np.concatenate(cs).astype(np.int32)).to(dev)


def timed(fn, warm=2, reps=5):
    ts = []
    for k in range(warm + reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if k >= warm:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('what', nargs='?', default='det')
    ap.add_argument('--top', type=int, default=25)
    ap.add_argument('--cprofile', action='store_true')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    if a.what == 'det':
        from link_b200.scn import SpMiddleResNetFHDELKv3
        feats, idx = det_inputs(dev)
        net = SpMiddleResNetFHDELKv3(num_input_features=5, ds_factor=8).to(dev)

        def fwd():
            with torch.no_grad():
                return net(feats, idx, 1, [1440, 1440, 40])[0]

        def step():
            f = feats.clone().requires_grad_(True)
            net(f, idx, 1, [1440, 1440, 40])[0].square().mean().backward()
            net.zero_grad(set_to_none=True)
        net.eval()
        print(f'det backbone N={len(idx)}: fwd (eval, fused) {timed(fwd):.2f} ms', flush=True)
        net.train()
    else:
        from link_b200 import SparseTensor
        from link_b200.linkencoder import ELKEncoder
        feats, coords = enc_inputs(dev)
        net = ELKEncoder(num_classes=19, cr=1.0, baseop='cos', r=3, s=7, groups=2).to(dev)
        target = torch.randint(0, 19, (coords.shape[0],), device=dev)

        def fwd():
            with torch.no_grad():
                return net(SparseTensor(feats, coords, 1))

        def fwd16():
            with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
                return net(SparseTensor(feats.to(torch.bfloat16), coords, 1))

        def step():
            loss = torch.nn.functional.cross_entropy(net(SparseTensor(feats, coords, 1)), target)
            loss.backward()
            net.zero_grad(set_to_none=True)
        net.eval()
        print(f'encoder N={coords.shape[0]}: fwd (eval, fused) {timed(fwd):.2f} ms', flush=True)
        print(f'encoder N={coords.shape[0]}: fwd (eval, fused, bf16 activations) {timed(fwd16):.2f} ms', flush=True)
        net.train()
    print(f'fwd+bwd (train): {timed(step):.2f} ms', flush=True)
    if a.cprofile:
        # host side: where the python time of a training step goes (the step is host-bound)
        import cProfile
        import pstats
        import time
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            step()
        host_ms = (time.perf_counter() - t0) * 1e3 / 3
        torch.cuda.synchronize()
        print(f'host enqueue per step: {host_ms:.2f} ms', flush=True)
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(3):
            step()
        pr.disable()
        torch.cuda.synchronize()
        pstats.Stats(pr).sort_stats('tottime').print_stats(45)
        pstats.Stats(pr).sort_stats('cumulative').print_stats(40)
        return
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=a.top, max_name_column_width=70))


if __name__ == '__main__':
    main()
