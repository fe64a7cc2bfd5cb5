"""Rotated BEV IoU / rotated NMS against the reference's own CPU implementation
(tests/golden/iou3d.npz from tests/golden/make_iou3d_golden.py; det3d/ops/iou3d_nms/src/iou3d_cpu.cpp
compiled unmodified).  Pure PyTorch: runs on CPU."""
import os

import numpy as np
import pytest
import torch

from link_b200.iou3d import boxes_iou_bev, nms_fixed_point, rotate_nms, rotate_nms_pcdet

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'iou3d.npz')


@pytest.fixture(scope='module')
def gold():
    return dict(np.load(GOLD))


def test_boxes_iou_bev_matches_reference(gold):
    got = boxes_iou_bev(torch.from_numpy(gold['a']), torch.from_numpy(gold['b']), block=64).numpy()
    np.testing.assert_allclose(got, gold['iou'], rtol=1e-4, atol=2e-5)
    assert ((got > 0) == (gold['iou'] > 0)).mean() > 0.999       # same overlap / no-overlap decisions


def test_known_values():
    import math
    b = torch.tensor([[0, 0, 0, 2, 1, 1, 0.0], [0, 0, 0, 2, 1, 1, math.pi / 2], [9, 9, 0, 1, 1, 1, 0.3]])
    iou = boxes_iou_bev(b, b)
    assert torch.allclose(iou.diagonal(), torch.ones(3), atol=1e-5)
    assert abs(float(iou[0, 1]) - 1 / 3) < 1e-5 and float(iou[0, 2]) == 0.0
    assert boxes_iou_bev(b[:0], b).shape == (0, 3)


@pytest.mark.parametrize('tag', ['t02', 't05', 't001'])
def test_rotate_nms_pcdet_matches_reference(gold, tag):
    thresh, pre, post = gold[f'nms_{tag}_cfg']
    sel = rotate_nms_pcdet(torch.from_numpy(gold['nms_boxes']), torch.from_numpy(gold['nms_scores']), float(thresh),
                           pre_maxsize=None if pre < 0 else int(pre), post_max_size=None if post < 0 else int(post))
    assert sel.tolist() == gold[f'nms_{tag}'].tolist()


def test_nms_fixed_point_equals_sequential_greedy():
    rng = np.random.default_rng(0)
    for n, p in ((1, 0.5), (50, 0.3), (200, 0.05), (120, 0.9)):
        m = rng.random((n, n)) < p
        keep = nms_fixed_point(torch.from_numpy(m)).numpy()
        supp, want = np.zeros(n, bool), np.zeros(n, bool)
        for i in range(n):
            if not supp[i]:
                want[i] = True
                supp |= m[i] & (np.arange(n) > i)
        assert np.array_equal(keep, want)
    chain = torch.zeros(64, 64, dtype=torch.bool)
    chain[torch.arange(63), torch.arange(1, 64)] = True            # i suppresses i+1: the longest possible chain
    assert nms_fixed_point(chain).tolist() == [i % 2 == 0 for i in range(64)]
    assert rotate_nms(torch.zeros(0, 7), torch.zeros(0), 0.5).numel() == 0


def test_kernel_arithmetic_on_host_matches_reference(gold):
    """csrc/iou3d.cu evaluates pairs with a __host__ __device__ function; the host instantiation
    (lk_boxes_iou_bev_hostcheck, a test hook) is held to the reference fixture here, the device
    instantiation in the GPU suite."""
    from link_b200 import _capi
    a = np.ascontiguousarray(gold['a'], np.float32)
    b = np.ascontiguousarray(gold['b'], np.float32)
    out = np.zeros((len(a), len(b)), np.float32)
    _capi.check(_capi.lib().lk_boxes_iou_bev_hostcheck(a.ctypes.data, len(a), b.ctypes.data, len(b), out.ctypes.data), 'hostcheck')
    np.testing.assert_allclose(out, gold['iou'], rtol=1e-5, atol=2e-6)
    assert np.array_equal(out > 0, gold['iou'] > 0)
