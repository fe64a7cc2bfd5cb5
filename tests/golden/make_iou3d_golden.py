"""Golden fixture for the rotated BEV IoU / rotated NMS (tests/golden/iou3d.npz): the reference's own
CPU implementation (det3d/ops/iou3d_nms/src/iou3d_cpu.cpp, compiled unmodified by
oracle/build_ref.py::build_iou3d) on seeded boxes, plus the greedy NMS the reference's nms_gpu
performs (keep the best box, drop every later box with IoU > thresh) evaluated on that IoU matrix.
Build container only."""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def boxes(seed, n, spread):
    g = torch.Generator().manual_seed(seed)
    xy = (torch.rand(n, 2, generator=g) - 0.5) * spread
    z = torch.randn(n, 1, generator=g)
    dims = torch.rand(n, 3, generator=g) * torch.tensor([4.0, 2.0, 1.5]) + torch.tensor([0.5, 0.4, 0.5])
    yaw = (torch.rand(n, 1, generator=g) - 0.5) * 4 * math.pi
    return torch.cat([xy, z, dims, yaw], 1).contiguous()


def greedy(iou, thresh):
    keep, supp = [], np.zeros(len(iou), bool)
    for i in range(len(iou)):
        if supp[i]:
            continue
        keep.append(i)
        supp |= (iou[i] > thresh) & (np.arange(len(iou)) > i)
    return np.array(keep)


def main():
    from oracle import build_ref
    ref = build_ref.load_iou3d()
    out = {}
    a, b = boxes(0, 150, 30.0), boxes(1, 120, 30.0)
    special = torch.tensor([[0, 0, 0, 2, 1, 1, 0.0], [0, 0, 0, 2, 1, 1, math.pi / 2], [0.5, 0.25, 0, 2, 1, 1, 0.0],
                            [0, 0, 0, 2, 1, 1, 0.3], [10, 10, 0, 1, 1, 1, 0.7], [0, 0, 0, 4, 4, 1, 1.0],
                            [0.1, 0.1, 0, 0.5, 0.5, 1, 2.0]], dtype=torch.float32)
    a = torch.cat([a, special]).contiguous()
    iou = torch.zeros(len(a), len(b) + len(special))
    bb = torch.cat([b, special]).contiguous()
    ref.boxes_iou_bev_cpu(a, bb, iou)
    out['a'], out['b'], out['iou'] = a.numpy(), bb.numpy(), iou.numpy()
    # NMS: dense cluster so that suppression chains exist; det3d convention boxes (x, y, z, w, l, h, yaw)
    d = boxes(2, 400, 25.0)
    g = torch.Generator().manual_seed(3)
    scores = torch.rand(400, generator=g)
    conv = d[:, [0, 1, 2, 4, 3, 5, 6]].clone()
    conv[:, 6] = -conv[:, 6] - math.pi / 2
    for tag, thresh, pre, post in (('t02', 0.2, 300, 83), ('t05', 0.5, None, None), ('t001', 0.01, 1000, 50)):
        order = torch.argsort(scores, descending=True)
        if pre is not None:
            order = order[:pre]
        bs = conv[order].contiguous()
        m = torch.zeros(len(bs), len(bs))
        ref.boxes_iou_bev_cpu(bs, bs, m)
        keep = order.numpy()[greedy(m.numpy(), thresh)]
        out[f'nms_{tag}'] = keep[:post] if post is not None else keep
        out[f'nms_{tag}_cfg'] = np.array([thresh, -1 if pre is None else pre, -1 if post is None else post])
        print(tag, len(out[f'nms_{tag}']))
    out['nms_boxes'], out['nms_scores'] = d.numpy(), scores.numpy()
    np.savez_compressed(os.path.join(HERE, 'iou3d.npz'), **out)
    print('iou', iou.shape, 'nonzero', int((iou > 0).sum()), 'max', float(iou.max()))


if __name__ == '__main__':
    main()
