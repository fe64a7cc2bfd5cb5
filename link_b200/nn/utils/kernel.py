from functools import lru_cache
from typing import Tuple, Union

import numpy as np
import torch

from link_b200.utils import make_ntuple

__all__ = ['get_kernel_offsets']


@lru_cache(maxsize=256)
def _offsets_np(size: Tuple[int, ...], stride: Tuple[int, ...], dilation: Tuple[int, ...]):
    axes = [np.arange(-size[k] // 2 + 1, size[k] // 2 + 1) * stride[k] * dilation[k]
            for k in range(3)]
    # Ordering rule of the reference (torchsparse/nn/utils/kernel.py:24-29): odd kernel volume
    # -> x fastest (MinkowskiEngine weight layout); even volume -> z fastest.
    if int(np.prod(size)) % 2 == 1:
        grid = [[x, y, z] for z in axes[2] for y in axes[1] for x in axes[0]]
    else:
        grid = [[x, y, z] for x in axes[0] for y in axes[1] for z in axes[2]]
    return np.asarray(grid, dtype=np.int32).reshape(-1, 3)


_device_cache = {}


def get_kernel_offsets(size: Union[int, Tuple[int, ...]], stride: Union[int, Tuple[int, ...]] = 1,
                       dilation: Union[int, Tuple[int, ...]] = 1, device='cpu') -> torch.Tensor:
    """int32 [K, 3] offset table; same values and ORDER as the reference.  Tables are cached per
    (size, stride, dilation, device) so the hot path never re-uploads them (the reference builds
    a numpy array and copies it host->device on every call)."""
    key = (make_ntuple(size, 3), make_ntuple(stride, 3), make_ntuple(dilation, 3), str(device))
    t = _device_cache.get(key)
    if t is None:
        t = torch.from_numpy(_offsets_np(*key[:3]).copy()).to(device)
        _device_cache[key] = t
    return t
