from typing import Optional

import torch

from link_b200 import _capi

__all__ = ['sphash']


def sphash(coords: torch.Tensor, offsets: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Same signature and bit-exact values as the reference `F.sphash`
    (torchsparse/nn/functional/hash.py:10-37): int32 [N,4] -> int64 [N], or with int32 [K,3]
    offsets -> int64 [K,N]."""
    assert coords.dtype == torch.int, coords.dtype
    assert coords.ndim == 2 and coords.shape[1] == 4, coords.shape
    coords = coords.contiguous()
    n = coords.shape[0]
    L = _capi.lib()
    if offsets is None:
        out = torch.empty(n, dtype=torch.int64, device=coords.device)
        _capi.check(L.lk_hash(_capi.ptr(coords), n, _capi.ptr(out), _capi.stream()), 'lk_hash')
        return out
    assert offsets.dtype == torch.int, offsets.dtype
    assert offsets.ndim == 2 and offsets.shape[1] == 3, offsets.shape
    offsets = offsets.contiguous().to(coords.device)
    k = offsets.shape[0]
    out = torch.empty(k, n, dtype=torch.int64, device=coords.device)
    _capi.check(L.lk_kernel_hash(_capi.ptr(coords), n, _capi.ptr(offsets), k, _capi.ptr(out),
                                 _capi.stream()), 'lk_kernel_hash')
    return out
