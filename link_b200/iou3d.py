"""Rotated bird's-eye-view IoU and rotated NMS on the device, as batched tensor
operations -- no per-box loop, no host round trip per box.

Reference: detection/det3d/ops/iou3d_nms (CUDA kernels iou3d_nms_kernel.cu:236-414 and their CPU twin
src/iou3d_cpu.cpp:59-252), called through `rotate_nms_pcdet`
(detection/det3d/core/bbox/box_torch_ops.py:248-277).  The arithmetic follows the reference's
`box_overlap`: the intersection polygon of two rotated rectangles is the set of proper edge
crossings plus the corners of either box that lie inside the other (margin 1e-2), ordered by angle
around their mean and measured with the shoelace sum.

CUDA tensors take the fused kernels of csrc/iou3d.cu (lk_boxes_iou_bev: one thread per pair;
lk_nms_bev: overlap bitmasks + an on-device greedy scan); the tensor-op formulation below serves CPU
tensors (fixture generation and the CPU tests) and stays the readable statement of the arithmetic.
Parity: tests/test_iou3d_cpu.py against the reference's own compiled
`boxes_iou_bev_cpu` (oracle/_ref/iou3d_cpu_ref.so in the build container, tests/golden/iou3d.npz
everywhere).
"""
import math

import torch

__all__ = ['boxes_iou_bev', 'nms_fixed_point', 'rotate_nms', 'rotate_nms_pcdet']

_EPS = 1e-8
_MARGIN = 1e-2


def _corners(boxes: torch.Tensor) -> torch.Tensor:
    """[K, 7] (x, y, z, dx, dy, dz, heading) -> [K, 4, 2] corners in the reference's order
    (iou3d_cpu.cpp:135-161): (-,-), (+,-), (+,+), (-,+) rotated by `heading` about the centre."""
    hx, hy = boxes[:, 3] / 2, boxes[:, 4] / 2
    sx = torch.stack([-hx, hx, hx, -hx], dim=1)
    sy = torch.stack([-hy, -hy, hy, hy], dim=1)
    c, s = torch.cos(boxes[:, 6])[:, None], torch.sin(boxes[:, 6])[:, None]
    x = sx * c - sy * s + boxes[:, 0:1]
    y = sx * s + sy * c + boxes[:, 1:2]
    return torch.stack([x, y], dim=2)


def _cross3(p1, p2, p0):
    return (p1[..., 0] - p0[..., 0]) * (p2[..., 1] - p0[..., 1]) - (p2[..., 0] - p0[..., 0]) * (p1[..., 1] - p0[..., 1])


def _inside(box, pts):
    """pts [..., P, 2] inside box [..., 7] (broadcast over leading dims), iou3d_cpu.cpp:75-85."""
    c, s = torch.cos(-box[..., 6])[..., None], torch.sin(-box[..., 6])[..., None]
    dx, dy = pts[..., 0] - box[..., 0:1], pts[..., 1] - box[..., 1:2]
    rx = dx * c - dy * s
    ry = dx * s + dy * c
    return (rx.abs() < box[..., 3:4] / 2 + _MARGIN) & (ry.abs() < box[..., 4:5] / 2 + _MARGIN)


def _overlap_block(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Intersection areas [Na, Nb] of rotated rectangles a [Na, 7], b [Nb, 7]."""
    na, nb = a.shape[0], b.shape[0]
    ca, cb = _corners(a), _corners(b)                                  # [Na,4,2], [Nb,4,2]
    # edge i of a: p0 = ca[i] -> p1 = ca[i+1]; edge j of b: q0 = cb[j] -> q1 = cb[j+1]
    p0 = ca[:, None, :, None, :].expand(na, nb, 4, 4, 2)
    p1 = ca.roll(-1, dims=1)[:, None, :, None, :].expand(na, nb, 4, 4, 2)
    q0 = cb[None, :, None, :, :].expand(na, nb, 4, 4, 2)
    q1 = cb.roll(-1, dims=1)[None, :, None, :, :].expand(na, nb, 4, 4, 2)
    s1, s2 = _cross3(q0, p1, p0), _cross3(p1, q1, p0)
    s3, s4 = _cross3(p0, q1, q0), _cross3(q1, p1, q0)
    hit = (s1 * s2 > 0) & (s3 * s4 > 0)                                # proper crossing (iou3d_cpu.cpp:91-97)
    s5 = _cross3(q1, p1, p0)
    den = s5 - s1
    den_ok = den.abs() > _EPS
    den_s = torch.where(den_ok, den, torch.ones_like(den))
    ix = (s5 * q0[..., 0] - s1 * q1[..., 0]) / den_s
    iy = (s5 * q0[..., 1] - s1 * q1[..., 1]) / den_s
    # nearly parallel supporting lines: general line-line intersection (iou3d_cpu.cpp:106-113)
    a0, b0 = p0[..., 1] - p1[..., 1], p1[..., 0] - p0[..., 0]
    c0 = p0[..., 0] * p1[..., 1] - p1[..., 0] * p0[..., 1]
    a1, b1 = q0[..., 1] - q1[..., 1], q1[..., 0] - q0[..., 0]
    c1 = q0[..., 0] * q1[..., 1] - q1[..., 0] * q0[..., 1]
    d = a0 * b1 - a1 * b0
    d_s = torch.where(d == 0, torch.ones_like(d), d)
    ix = torch.where(den_ok, ix, (b0 * c1 - b1 * c0) / d_s)
    iy = torch.where(den_ok, iy, (a1 * c0 - a0 * c1) / d_s)
    cross_pts = torch.stack([ix, iy], dim=-1).reshape(na, nb, 16, 2)
    cross_ok = hit.reshape(na, nb, 16)
    # corners of b inside a, corners of a inside b
    b_pts = cb[None].expand(na, nb, 4, 2)
    a_pts = ca[:, None].expand(na, nb, 4, 2)
    b_in = _inside(a[:, None, :], b_pts)
    a_in = _inside(b[None, :, :], a_pts)
    pts = torch.cat([cross_pts, b_pts, a_pts], dim=2)                  # [Na,Nb,24,2]
    ok = torch.cat([cross_ok, b_in, a_in], dim=2)
    cnt = ok.sum(dim=2)
    w = ok.unsqueeze(-1).to(pts.dtype)
    centre = (pts * w).sum(dim=2) / cnt.clamp(min=1).unsqueeze(-1).to(pts.dtype)
    ang = torch.atan2(pts[..., 1] - centre[:, :, None, 1], pts[..., 0] - centre[:, :, None, 0])
    ang = torch.where(ok, ang, torch.full_like(ang, math.inf))        # invalid candidates sort last
    order = torch.argsort(ang, dim=2, stable=True)
    pts = torch.gather(pts, 2, order.unsqueeze(-1).expand(-1, -1, -1, 2))
    ok = torch.gather(ok, 2, order)
    first = pts[:, :, 0:1, :]
    rel = torch.where(ok.unsqueeze(-1), pts - first, torch.zeros_like(pts))   # invalid -> the fan's apex: zero area
    area = (rel[:, :, :-1, 0] * rel[:, :, 1:, 1] - rel[:, :, :-1, 1] * rel[:, :, 1:, 0]).sum(dim=2)
    return torch.where(cnt > 0, area.abs() / 2, torch.zeros_like(area))


def boxes_iou_bev(boxes_a: torch.Tensor, boxes_b: torch.Tensor, block: int = 256) -> torch.Tensor:
    """Rotated BEV IoU [N, M] of boxes `(x, y, z, dx, dy, dz, heading)` (iou3d_cpu.cpp:222-252),
    evaluated in row blocks to bound the [rows, M, 24] intermediates."""
    a, b = boxes_a.float(), boxes_b.float()
    out = torch.empty(a.shape[0], b.shape[0], dtype=torch.float32, device=a.device)
    if a.shape[0] == 0 or b.shape[0] == 0:
        return out
    if a.is_cuda:                                   # fused kernel: one thread per pair (csrc/iou3d.cu)
        from link_b200 import _capi
        a, b = a.contiguous(), b.contiguous()
        _capi.check(_capi.lib().lk_boxes_iou_bev(_capi.ptr(a), a.shape[0], _capi.ptr(b), b.shape[0], _capi.ptr(out),
                                                 _capi.stream()), 'lk_boxes_iou_bev')
        return out
    area_b = (b[:, 3] * b[:, 4])[None, :]
    for r0 in range(0, a.shape[0], block):
        blk = a[r0:r0 + block]
        inter = _overlap_block(blk, b)
        union = (blk[:, 3] * blk[:, 4])[:, None] + area_b - inter
        out[r0:r0 + block] = inter / union.clamp(min=_EPS)
    return out


def nms_fixed_point(suppresses: torch.Tensor) -> torch.Tensor:
    """Greedy NMS over boxes already sorted by descending score, given the boolean matrix
    `suppresses[i, j]` (box i removes box j when i is kept; only i < j is used).  `keep[j] = not
    any_{i<j}(keep[i] and suppresses[i, j])` has a unique solution; re-evaluating the right-hand
    side from keep = all reaches it after as many sweeps as the longest suppression chain.
    Returns the boolean keep mask."""
    n = suppresses.shape[0]
    upper = torch.triu(suppresses, diagonal=1)
    keep = torch.ones(n, dtype=torch.bool, device=suppresses.device)
    for _ in range(n):
        new_keep = ~(upper & keep[:, None]).any(dim=0)
        if bool((new_keep == keep).all()):
            break
        keep = new_keep
    return keep


def rotate_nms(boxes: torch.Tensor, scores: torch.Tensor, thresh: float, pre_maxsize=None,
               post_max_size=None) -> torch.Tensor:
    """Greedy rotated NMS (iou3d_nms.cpp nms_gpu semantics: a box is dropped when a kept,
    higher-scoring box overlaps it with BEV IoU > thresh).  boxes `(x, y, z, dx, dy, dz, heading)`.
    Returns indices into the input, highest score first."""
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    if order.numel() == 0:
        return order
    b = boxes[order].contiguous()
    if b.is_cuda and b.shape[0] <= 65536:
        # device NMS (csrc/iou3d.cu): 64 x 64 overlap bitmasks + a one-warp greedy scan, no host round trip
        from link_b200 import _capi
        b = b.float().contiguous()
        n = b.shape[0]
        L = _capi.lib()
        ws_bytes = L.lk_nms_bev_ws_bytes(n)
        ws = torch.empty(ws_bytes // 8 + 1, dtype=torch.int64, device=b.device)
        keep8 = torch.empty(n, dtype=torch.uint8, device=b.device)
        _capi.check(L.lk_nms_bev(_capi.ptr(b), n, float(thresh), _capi.ptr(ws), ws.numel() * 8, _capi.ptr(keep8),
                                 _capi.stream()), 'lk_nms_bev')
        keep = keep8.bool()
    else:
        keep = nms_fixed_point(boxes_iou_bev(b, b) > thresh)
    sel = order[keep]
    return sel[:post_max_size] if post_max_size is not None else sel


def rotate_nms_pcdet(boxes: torch.Tensor, scores: torch.Tensor, thresh: float, pre_maxsize=None,
                     post_max_size=None) -> torch.Tensor:
    """det3d's entry point (box_torch_ops.py:248-277): boxes `(x, y, z, w, l, h, yaw)` in det3d's
    convention are converted to the IoU op's `(x, y, z, l, w, h, -yaw - pi/2)` first."""
    b = boxes[:, [0, 1, 2, 4, 3, 5, -1]].clone()
    b[:, -1] = -b[:, -1] - math.pi / 2
    return rotate_nms(b, scores, thresh, pre_maxsize, post_max_size)
