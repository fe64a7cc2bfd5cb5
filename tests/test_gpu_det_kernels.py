"""GPU tests of the SURVEY 8f custom kernels: the fused sparse -> dense BEV scatter (+ its backward) and
the on-device rotated NMS (overlap bitmasks + greedy scan)."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'gpu-marked tests need a CUDA device'
    from link_b200 import _capi
    _capi.lib()
    return torch.device('cuda:0')


@pytest.mark.parametrize('n,c,shape', [(5000, 128, (2, 2, 180, 180)), (777, 16, (1, 5, 40, 33)), (1, 64, (3, 1, 7, 9)),
                                       (0, 32, (1, 2, 8, 8))])
def test_bev_scatter_matches_dense_and_backward(dev, n, c, shape):
    """SparseConvTensor.dense() on the fused kernel == spconv's zero-fill + scatter + permute
    (scn.py:612-617) bit for bit; its backward == the gather of the same positions."""
    from link_b200.ts_elk import SparseConvTensor
    B, D, H, W = shape
    gen = torch.Generator().manual_seed(n + c)
    flat = torch.randperm(B * D * H * W, generator=gen)[:n]
    idx = torch.stack([flat // (D * H * W), (flat // (H * W)) % D, (flat // W) % H, flat % W], 1).int().to(dev)
    feats = torch.randn(n, c, generator=gen).to(dev).requires_grad_(True)
    got = SparseConvTensor(feats, idx, [D, H, W], B).dense()
    ref_in = feats.detach().clone().requires_grad_(True)
    want = torch.zeros(B, D, H, W, c, device=dev)
    i = idx.long()
    want = want.index_put((i[:, 0], i[:, 1], i[:, 2], i[:, 3]), ref_in).permute(0, 4, 1, 2, 3).contiguous()
    assert got.shape == (B, c, D, H, W) and torch.equal(got, want)
    go = torch.randn(got.shape, generator=gen).to(dev)
    got.backward(go)
    want.backward(go)
    assert torch.equal(feats.grad, ref_in.grad)


def _random_boxes(n, seed, extent=40.0):
    g = torch.Generator().manual_seed(seed)
    xy = (torch.rand(n, 2, generator=g) - 0.5) * extent
    z = torch.zeros(n, 1)
    dims = torch.rand(n, 3, generator=g) * 4 + 0.5
    yaw = (torch.rand(n, 1, generator=g) - 0.5) * 2 * math.pi
    return torch.cat([xy, z, dims, yaw], 1), torch.rand(n, generator=g)


@pytest.mark.parametrize('n,thresh', [(1500, 0.2), (64, 0.5), (65, 0.1), (4000, 0.7), (1, 0.2)])
def test_device_nms_equals_fixed_point_formulation(dev, n, thresh):
    """lk_nms_bev (bitmask + on-device greedy scan) selects exactly what greedy NMS over the pairwise IoU
    matrix of the SAME kernel selects (nms_fixed_point = the reference's sequential rule)."""
    from link_b200.iou3d import boxes_iou_bev, nms_fixed_point, rotate_nms
    boxes, scores = _random_boxes(n, seed=n)
    boxes, scores = boxes.to(dev), scores.to(dev)
    sel = rotate_nms(boxes, scores, thresh)
    order = scores.sort(0, descending=True)[1]
    b = boxes[order].contiguous()
    want = order[nms_fixed_point(boxes_iou_bev(b, b) > thresh)]
    assert torch.equal(sel, want)
    # cross-check the rule itself with a plain sequential loop on the host
    iou = boxes_iou_bev(b, b).cpu().numpy()
    keep, alive = [], np.ones(n, bool)
    for i in range(n):
        if alive[i]:
            keep.append(i)
            alive &= ~(iou[i] > thresh) | (np.arange(n) <= i)
    assert sel.tolist() == order.cpu().numpy()[keep].tolist()


def test_device_nms_vs_reference_fixture_misses_are_threshold_noise(dev):
    """Against the selection generated from the reference's own CPU IoU (tests/golden/iou3d.npz): every
    box on which the two selections differ owes it to an IoU within 1e-4 of the threshold somewhere in its
    suppression chain -- checked by re-running the greedy rule with the fixture's IoU matrix perturbed
    only inside that band."""
    from link_b200.iou3d import boxes_iou_bev, rotate_nms_pcdet
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'iou3d.npz')))
    boxes, scores = torch.from_numpy(g['nms_boxes']).to(dev), torch.from_numpy(g['nms_scores']).to(dev)
    sel = rotate_nms_pcdet(boxes, scores, 0.2, pre_maxsize=300, post_max_size=83).cpu().numpy()
    want = g['nms_t02']
    if np.array_equal(sel, want):
        return
    # our IoU matrix of the sorted candidates vs the threshold: entries that decide differently than a
    # matrix 1e-4 away would are the only admissible cause of a difference
    order = scores.sort(0, descending=True)[1][:300]
    b = boxes[order][:, [0, 1, 2, 4, 3, 5, 6]].clone()
    b[:, -1] = -b[:, -1] - math.pi / 2
    iou = boxes_iou_bev(b.contiguous(), b.contiguous()).cpu().numpy()
    band = np.abs(iou - 0.2) < 1e-4
    assert band.any(), 'selections differ although no IoU lies within 1e-4 of the threshold'

    def greedy(m):
        n = len(m)
        keep, alive = [], np.ones(n, bool)
        for i in range(n):
            if alive[i]:
                keep.append(i)
                alive &= ~m[i] | (np.arange(n) <= i)
        return order.cpu().numpy()[keep][:83]
    lo, hi = greedy(iou > 0.2 + 1e-4), greedy(iou > 0.2 - 1e-4)
    assert np.array_equal(want, lo) or np.array_equal(want, hi) or np.array_equal(want, greedy(np.where(band, ~(iou > 0.2), iou > 0.2)))


@pytest.mark.parametrize('n,thresh', [(3000, 4.0), (64, 0.5), (129, 100.0), (1, 1.0)])
def test_device_circle_nms_equals_sequential_rule(dev, n, thresh):
    """lk_nms_circle == the reference's sequential loop (circle_nms_jit.py:4-28) == the masked-reduction form."""
    from link_b200.centerpoint import circle_nms
    g = torch.Generator().manual_seed(n)
    centers = (torch.rand(n, 2, generator=g) * 60 - 30)
    scores = torch.rand(n, generator=g)
    sel = circle_nms(centers.to(dev), scores.to(dev), thresh, post_max_size=10 ** 6)
    order = torch.argsort(scores, descending=True).numpy()
    c = centers.numpy()[order]
    keep, dead = [], np.zeros(n, bool)
    for i in range(n):
        if dead[i]:
            continue
        keep.append(i)
        d = ((c[i] - c) ** 2).sum(1)
        dead |= (d <= thresh) & (np.arange(n) > i)
    assert sel.cpu().numpy().tolist() == order[keep].tolist()
    assert circle_nms(centers, scores, thresh, post_max_size=10 ** 6).tolist() == order[keep].tolist()   # CPU tensor path
