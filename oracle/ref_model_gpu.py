"""Reference GPU arm, unmodified end to end: the reference's OWN python package and model classes (staged
untouched under baseline/_ref/py by oracle/build_ref.py::stage_python) on the reference's OWN CUDA
kernels (oracle/_ref/backend_cuda.so, compiled from /root/reference for sm_100a).  Nothing of link_b200 is
on this path except the synthetic-scan generator.  TEST / BENCH INFRASTRUCTURE ONLY; run as a script
(own process: it registers the reference's package as `torchsparse`):

    python oracle/ref_model_gpu.py [--voxels 120000] [--steps 5] [--warmup 2] [--what encoder|block]

Prints one JSON line {what, n, ms_per_step, voxels_per_s, steps}."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--voxels', type=int, default=120_000)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--what', default='encoder', choices=['encoder', 'block'])
    a = ap.parse_args()
    import numpy as np
    import torch
    from oracle import build_ref
    stage = os.path.join(ROOT, 'baseline', '_ref', 'py')
    if not (os.path.isdir(os.path.join(stage, 'torchsparse')) and os.path.exists(build_ref.OUT_CUDA)):
        print(json.dumps({'unavailable': 'baseline/_ref/py or oracle/_ref/backend_cuda.so not built'}))
        return
    backend = build_ref.load_cuda()
    sys.path.insert(0, stage)
    sys.modules['torchsparse.backend'] = backend
    import torchsparse
    torchsparse.backend = backend
    from core.models.semantic_kitti import linkencoder as ref_models
    from link_b200.utils.synthetic import kitti_like_voxels      # the scan generator only
    dev = torch.device('cuda:0')
    c3, f4 = kitti_like_voxels(a.voxels, seed=0)
    coords = torch.from_numpy(np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32)).to(dev)
    torch.manual_seed(0)
    if a.what == 'encoder':
        net = ref_models.ELKEncoder(num_classes=19, cr=1.0, baseop='cos', r=3, s=7, groups=2).to(dev).eval()
        feats = torch.from_numpy(f4.astype(np.float32)).to(dev)
        run = lambda: net(torchsparse.SparseTensor(feats, coords, 1))
    else:
        net = ref_models.ELKBlock(64, 64, groups=2, baseop='cos').to(dev).eval()
        feats = torch.randn(coords.shape[0], 64, device=dev)
        run = lambda: net(torchsparse.SparseTensor(feats.clone(), coords, 1), 7, 3)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ms = []
    with torch.no_grad():
        for k in range(a.warmup + a.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            if k >= a.warmup:
                ms.append(e0.elapsed_time(e1))
    med = float(np.median(ms))
    print(json.dumps({'what': a.what, 'n': int(coords.shape[0]), 'ms_per_step': med, 'voxels_per_s': coords.shape[0] / (med * 1e-3),
                      'steps': a.steps, 'kind': 'reference python package + model class (baseline/_ref/py, untouched) on the '
                                                'reference CUDA backend (oracle/_ref/backend_cuda.so), L2 flushed between '
                                                'steps, CUDA events, median'}))


if __name__ == '__main__':
    main()
