import torch

from link_b200 import _capi

__all__ = ['spcount']


def spcount(coords: torch.Tensor, num) -> torch.Tensor:
    """Reference `F.spcount` (torchsparse/nn/functional/count.py:8-16): int32 histogram of the
    non-negative entries of an int32 index vector into `num` bins."""
    coords = coords.contiguous()
    assert coords.dtype == torch.int, coords.dtype
    num = int(num)
    out = torch.empty(num, dtype=torch.int32, device=coords.device)
    _capi.check(_capi.lib().lk_count(_capi.ptr(coords), coords.shape[0], _capi.ptr(out), num,
                                     _capi.stream()), 'lk_count')
    return out
