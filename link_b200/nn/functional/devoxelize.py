import torch
from torch.autograd import Function

from link_b200 import _capi

__all__ = ['spdevoxelize', 'calc_ti_weights']


def calc_ti_weights(coords: torch.Tensor, idx_query: torch.Tensor, scale: float = 1) -> torch.Tensor:
    """Trilinear interpolation weights [8, N] (reference: devoxelize.py:10-48).  Only used by the
    point<->voxel branch of SPVCNN-style models; kept for API completeness."""
    with torch.no_grad():
        p = coords
        pf = torch.floor(coords / scale) * scale if scale != 1 else torch.floor(coords)
        pc = pf + scale
        x, y, z = (p[:, i].view(-1, 1) for i in range(3))
        xf, yf, zf = (pf[:, i].view(-1, 1).float() for i in range(3))
        xc, yc, zc = (pc[:, i].view(-1, 1).float() for i in range(3))
        ws = [(xc - x) * (yc - y) * (zc - z), (xc - x) * (yc - y) * (z - zf),
              (xc - x) * (y - yf) * (zc - z), (xc - x) * (y - yf) * (z - zf),
              (x - xf) * (yc - y) * (zc - z), (x - xf) * (yc - y) * (z - zf),
              (x - xf) * (y - yf) * (zc - z), (x - xf) * (y - yf) * (z - zf)]
        w = torch.cat(ws, dim=1).transpose(1, 0).contiguous()
        if scale != 1:
            w /= scale ** 3
        w[idx_query == -1] = 0
        w /= torch.sum(w, dim=0) + 1e-8
    return w


class DevoxelizeFunction(Function):
    """Weighted R-neighbour gather (reference: DevoxelizeFunction, devoxelize.py:51-98)."""

    @staticmethod
    def forward(ctx, feats: torch.Tensor, coords: torch.Tensor, weights: torch.Tensor, r: int):
        in_dtype = feats.dtype
        feats = feats.contiguous().float()
        coords = coords.contiguous().int()
        weights = weights.contiguous().float()
        N, R = coords.shape
        assert R == r ** 3 or r is None, (R, r)
        n, c = feats.shape
        out = torch.empty(N, c, dtype=torch.float32, device=feats.device)
        _capi.check(_capi.lib().lk_devoxelize_fwd(_capi.ptr(feats), _capi.ptr(coords),
                                                  _capi.ptr(weights), N, R, c, _capi.ptr(out),
                                                  _capi.stream()), 'lk_devoxelize_fwd')
        ctx.for_backwards = (coords, weights, n, in_dtype)
        return out.to(in_dtype)

    @staticmethod
    def backward(ctx, grad_output: torch.Tensor):
        coords, weights, n, in_dtype = ctx.for_backwards
        g = grad_output.contiguous().float()
        N, c = g.shape
        out = torch.empty(n, c, dtype=torch.float32, device=g.device)
        _capi.check(_capi.lib().lk_devoxelize_bwd(_capi.ptr(g), _capi.ptr(coords),
                                                  _capi.ptr(weights), N, coords.shape[1], c, n,
                                                  _capi.ptr(out), _capi.stream()),
                    'lk_devoxelize_bwd')
        return out.to(in_dtype), None, None, None


def spdevoxelize(feats: torch.Tensor, coords: torch.Tensor, weights: torch.Tensor, r=2) -> torch.Tensor:
    return DevoxelizeFunction.apply(feats, coords, weights, r)
