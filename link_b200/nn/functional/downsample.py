from typing import Dict, Optional, Tuple, Union

import torch

import ctypes as C

from link_b200 import _capi
from link_b200.nn.functional import _index
from link_b200.nn.functional._index import unique_coords
from link_b200.nn.utils import get_kernel_offsets
from link_b200.utils import make_ntuple

__all__ = ['spdownsample']


def spdownsample(coords: torch.Tensor, stride: Union[int, Tuple[int, ...]] = 2,
                 kernel_size: Union[int, Tuple[int, ...]] = 2,
                 tensor_stride: Union[int, Tuple[int, ...]] = 1,
                 cache: Optional[Dict] = None) -> torch.Tensor:
    """Output coordinates of a strided sparse conv, identical (values AND row order: sorted by
    batch, x, y, z) to the reference (torchsparse/nn/functional/downsample.py:11-51).  The
    reference floors the coordinates and runs torch.unique(dim=0); here the floor is folded into
    a packed (b, x, y, z) key that is radix-sorted and uniqued on device."""
    stride = make_ntuple(stride, ndim=3)
    kernel_size = make_ntuple(kernel_size, ndim=3)
    tensor_stride = make_ntuple(tensor_stride, ndim=3)
    sample_stride = [stride[k] * tensor_stride[k] for k in range(3)]
    coords = coords.contiguous()
    if all(stride[k] in [1, kernel_size[k]] for k in range(3)):
        # one library call (pack keys -> radix sort -> unique -> unpack), then the size read-back
        n = coords.shape[0]
        bounds = _index.coord_bounds(coords, cache)
        spec, bits = _index.make_keyspec(bounds, sample_stride, (3, 0, 1, 2), sample_stride)
        L = _capi.lib()
        out = torch.empty(n, 4, dtype=torch.int32, device=coords.device)
        num = torch.empty(1, dtype=torch.int32, device=coords.device)
        ws_bytes = L.lk_downsample_ws_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=coords.device)
        _capi.check(L.lk_downsample(_capi.ptr(coords, torch.int32), n, C.byref(spec), bits, out.data_ptr(),
                                    num.data_ptr(), ws.data_ptr(), ws_bytes, _capi.stream()), 'lk_downsample')
        out = out[:int(num.item())]
        _index.register_floored_bounds(cache, out, bounds, sample_stride)
        return out
    # general case (kernel != stride): expand by the kernel offsets, keep the aligned candidates
    offsets = get_kernel_offsets(kernel_size, tensor_stride, device=coords.device)
    kv = offsets.size(0)
    ss = torch.tensor(sample_stride, dtype=torch.int, device=coords.device).unsqueeze(0)
    cmin = torch.min(coords[:, :3], dim=0, keepdim=True).values
    x = coords[:, :3].unsqueeze(1).repeat(1, kv, 1) + offsets
    b = coords[:, 3:].repeat(1, kv)
    cand = torch.cat([x.view(-1, 3), b.view(-1, 1)], dim=1)
    mask = (cand[:, :3] % ss == 0) & (cand[:, :3] >= cmin)
    cand = cand[torch.all(mask, dim=1)].contiguous()
    out, _, _ = unique_coords(cand, sample_stride, (3, 0, 1, 2), sample_stride, None)
    return out
