import torch
from torch.autograd import Function

from link_b200 import _capi

__all__ = ['spvoxelize']


class VoxelizeFunction(Function):
    """Scatter-mean (reference: VoxelizeFunction, torchsparse/nn/functional/voxelize.py:10-51).
    fp32 compute; other float dtypes are converted at the boundary."""

    @staticmethod
    def forward(ctx, feats: torch.Tensor, coords: torch.Tensor, counts: torch.Tensor):
        in_dtype = feats.dtype
        feats = feats.contiguous().float()
        coords = coords.contiguous().int()
        counts = counts.contiguous().int()
        n, c = feats.shape
        m = counts.shape[0]
        out = torch.empty(m, c, dtype=torch.float32, device=feats.device)
        _capi.check(_capi.lib().lk_voxelize_fwd(_capi.ptr(feats), _capi.ptr(coords),
                                                _capi.ptr(counts), n, m, c, _capi.ptr(out),
                                                _capi.stream()), 'lk_voxelize_fwd')
        ctx.for_backwards = (coords, counts, n, in_dtype)
        return out.to(in_dtype)

    @staticmethod
    def backward(ctx, grad_output: torch.Tensor):
        coords, counts, n, in_dtype = ctx.for_backwards
        g = grad_output.contiguous().float()
        m, c = g.shape
        out = torch.empty(n, c, dtype=torch.float32, device=g.device)
        _capi.check(_capi.lib().lk_voxelize_bwd(_capi.ptr(g), _capi.ptr(coords), _capi.ptr(counts),
                                                n, m, c, _capi.ptr(out), _capi.stream()),
                    'lk_voxelize_bwd')
        return out.to(in_dtype), None, None


def spvoxelize(feats: torch.Tensor, coords: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    return VoxelizeFunction.apply(feats, coords, counts)
