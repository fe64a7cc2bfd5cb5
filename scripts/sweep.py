"""BASELINE config 5 (active-voxel sweep) and config 4 (detection backbone) on one GPU.

    python scripts/sweep.py [--out gpurun_out/sweep.json] [--quick] [--skip-det]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/sweep.py ...
        N replicas (one process per GPU, frames are independent: no data-path collective): every row then carries
        n_gpus = N, ms = max over the ranks, voxels_per_s = N * n / that time

ELKBlock forward (cos, groups=2), N in {10k..500k} x C in {16,64,128} x (r,s) in {(2,3),(3,5),(3,7)},
synthetic SemanticKITTI-shaped scans, index + kernel maps rebuilt every iteration, L2 flushed
between iterations, CUDA events, median of 10 after 3 warm-ups.  Then SpMiddleResNetFHDELKv3
(nuScenes-shaped grid 1440 x 1440 x 40, ~120k voxels, batch 1) forward and forward+backward.
Prints one JSON line per configuration and writes them all to --out."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from link_b200 import SparseTensor
from link_b200.elk import ELKBlock
from link_b200.nn.functional import _index
from link_b200.utils.synthetic import kitti_like_voxels, lidar_scan


def timed(fn, flush, warm=3, reps=10):
    """Median device time of fn(): the L2 flush in front of every timed call is repeated until it outlasts
    the host's enqueue time of one call, so the event window holds device work, not launch waiting
    (see bench.py / scripts/step_graph_probe.py)."""
    import time
    ts, host = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); flush.zero_(); flush.zero_(); e1.record(); torch.cuda.synchronize()
    flush_ms = e0.elapsed_time(e1) / 2
    n_flush = 1
    for k in range(warm + reps):
        for _ in range(n_flush):
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        fn()
        host.append((time.perf_counter() - t0) * 1e3)
        e1.record()
        torch.cuda.synchronize()
        n_flush = int(min(64, max(2, np.ceil(2.0 * np.median(host[-3:]) / flush_ms) + 1)))
        if k >= warm:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='gpurun_out/sweep.json')
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--skip-det', action='store_true')
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (('WORLD_SIZE', '1'), ('RANK', '0'), ('LOCAL_RANK', '0')))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('gloo')

    def over_ranks(ms):
        """max over the ranks of a per-rank time (gloo, CPU tensor); identity for one process"""
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def emit(row):
        row['ms'] = over_ranks(row['ms'])
        row['n_gpus'] = world
        row['voxels_per_s'] = world * row['n'] / (row['ms'] * 1e-3)
        if rank == 0:
            print(json.dumps(row), flush=True)
        rows.append(row)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    ns = [10_000, 50_000, 120_000] if a.quick else [10_000, 20_000, 50_000, 120_000, 250_000, 500_000]
    base = {}
    for n_t in ns:
        if n_t <= 125_000:
            c3, _ = kitti_like_voxels(n_t, seed=1)
        else:
            # larger clouds: several independent 125k scans side by side (x offset > scan extent), so the
            # local structure (voxels per block, neighbours per voxel) stays that of a LiDAR scan
            parts = []
            for k in range(n_t // 125_000):
                if k not in base:
                    base[k] = kitti_like_voxels(125_000, seed=2 + k)[0]
                parts.append(base[k] + np.array([[4000 * k, 0, 0]], np.int32))
            c3 = np.concatenate(parts, 0)
        coords_h = np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32)
        coords = torch.from_numpy(coords_h).to(dev)
        lo, hi = coords_h.min(0), coords_h.max(0)
        n = len(coords_h)
        for c in (16, 64, 128):
            torch.manual_seed(0)
            blk = ELKBlock(c, c, groups=2, baseop='cos').to(dev).eval()
            feats = torch.randn(n, c, device=dev)
            for r, s in ((2, 3), (3, 5), (3, 7)):
                def step():
                    st = SparseTensor(feats.clone(), coords, 1)
                    _index.set_coord_bounds(st.kmaps, lo, hi)
                    with torch.no_grad():
                        return blk(st, s, r).F
                ms = timed(step, flush)
                emit({'workload': 'ELKBlock fwd', 'n': n, 'c': c, 'r': r, 's': s, 'ms': ms})
    # ---- detection backbone (config 4): nuScenes-shaped grid
    try:
        if a.skip_det:
            raise KeyboardInterrupt
        from link_b200.scn import SpMiddleResNetFHDELKv3
        pts = np.concatenate([lidar_scan(seed=10 + k, beams=32, azimuths=1100) for k in range(10)], 0)
        rng = [-54, -54, -5, 54, 54, 3]
        vs = np.array([0.075, 0.075, 0.2])
        keep = np.all((pts[:, :3] >= rng[:3]) & (pts[:, :3] < rng[3:]), axis=1)
        ijk = np.floor((pts[keep, :3] - np.array(rng[:3])) / vs).astype(np.int32)      # (x, y, z)
        ijk = np.unique(ijk, axis=0)
        ijk = ijk[np.random.default_rng(0).permutation(len(ijk))[:120_000]]
        idx = np.concatenate([np.zeros((len(ijk), 1), np.int32), ijk[:, ::-1]], 1).astype(np.int32)  # (b, z, y, x)
        featsd = torch.randn(len(idx), 5, device=dev)
        torch.manual_seed(0)
        net = SpMiddleResNetFHDELKv3(num_input_features=5, ds_factor=8).to(dev).eval()
        idx_d = torch.from_numpy(idx).to(dev)

        def det_fwd():
            with torch.no_grad():
                return net(featsd, idx_d, 1, [1440, 1440, 40])[0]
        ms = timed(det_fwd, flush, warm=2, reps=5)
        emit({'workload': 'SpMiddleResNetFHDELKv3 fwd (eval, fused)', 'n': len(idx), 'ms': ms})
        net.train()

        def det_fwd_bwd():
            f = featsd.clone().requires_grad_(True)
            out = net(f, idx_d, 1, [1440, 1440, 40])[0]
            out.square().mean().backward()
            net.zero_grad(set_to_none=True)
        ms = timed(det_fwd_bwd, flush, warm=2, reps=5)
        emit({'workload': 'SpMiddleResNetFHDELKv3 fwd+bwd (train)', 'n': len(idx), 'ms': ms})
        # ---- full detector of config 4: reader -> backbone -> RPN -> CenterHead + losses, fwd+bwd
        from link_b200.centerpoint import NUSC_TASKS, build_nusc_centerpoint
        torch.manual_seed(0)
        det = build_nusc_centerpoint(backbone=net).to(dev).train()
        g = torch.Generator().manual_seed(1)
        n_pts = torch.randint(1, 11, (len(idx),), generator=g)
        voxels = (torch.randn(len(idx), 10, 5, generator=g) * (torch.arange(10)[None, :, None] < n_pts[:, None, None])).to(dev)
        example = {'voxels': voxels, 'num_points': n_pts.to(dev), 'coordinates': idx_d, 'batch_size': 1,
                   'shape': [np.array([1440, 1440, 40])], 'hm': [], 'ind': [], 'mask': [], 'cat': [], 'anno_box': []}
        for t in NUSC_TASKS:                       # CenterPoint targets: 180 x 180 heat maps, <= 500 objects
            k = len(t['class_names'])
            example['hm'].append((torch.rand(1, k, 180, 180, generator=g) ** 4).to(dev))
            example['ind'].append(torch.randint(0, 180 * 180, (1, 500), generator=g).to(dev))
            example['mask'].append((torch.rand(1, 500, generator=g) < 0.1).to(torch.uint8).to(dev))
            example['cat'].append(torch.randint(0, k, (1, 500), generator=g).to(dev))
            example['anno_box'].append(torch.randn(1, 500, 10, generator=g).to(dev))

        def detector_fwd_bwd():
            sum(det(example, return_loss=True)['loss']).backward()
            det.zero_grad(set_to_none=True)
        ms = timed(detector_fwd_bwd, flush, warm=2, reps=5)
        emit({'workload': 'CenterPoint detector (VFE + SpMiddleResNetFHDELKv3 + RPN + CenterHead + losses) fwd+bwd',
              'n': len(idx), 'ms': ms})
        # ---- inference: detector forward + CenterHead.predict (decode, filters, NMS on the device)
        det.eval()
        from link_b200.centerpoint import NUSC_TEST_CFG
        for name, circ in (('circle NMS', True), ('rotated NMS (lk_nms_bev)', False)):
            tcfg = dict(NUSC_TEST_CFG, circular_nms=circ, min_radius=[4, 12, 10, 1, 0.85, 0.175])

            def det_predict():
                det.test_cfg = tcfg
                with torch.no_grad():
                    return det(example, return_loss=False)
            ms = timed(det_predict, flush, warm=2, reps=5)
            emit({'workload': f'CenterPoint detector inference fwd + predict, {name}', 'n': len(idx), 'ms': ms})
    except KeyboardInterrupt:
        pass
    except Exception as e:   # the sweep above is the point; report rather than hide a failure here
        if rank == 0:
            print(json.dumps({'workload': 'detection backbone / detector', 'error': repr(e)[:300]}), flush=True)
    if rank == 0:
        os.makedirs(os.path.dirname(a.out) or '.', exist_ok=True)
        json.dump(rows, open(a.out, 'w'), indent=0)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
