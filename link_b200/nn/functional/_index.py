"""Host-side helpers around the device index primitives (key packing, radix sort/unique).

The only host<->device synchronisation of the whole index path is `coord_bounds` (one 8-int
read-back per input scan, cached in the tensor family's shared `kmaps` dict); callers that
need CUDA-graph capture can pre-seed the cache with `set_coord_bounds`."""
import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import torch

from link_b200 import _capi

BOUNDS_KEY = ('lk', 'bounds')


def coord_bounds(coords: torch.Tensor, cache: Optional[Dict] = None):
    """(lo[4], hi[4]) python ints of an int32 [N,4] coordinate tensor.  With `cache` (a
    SparseTensor.kmaps dict) the bounds of the family's finest level are reused: every derived
    level (floor to stride multiples, floor-div into blocks) stays inside them."""
    if cache is not None and BOUNDS_KEY in cache:
        return cache[BOUNDS_KEY]
    if coords.shape[0] == 0:
        b = ((0, 0, 0, 0), (0, 0, 0, 0))
    else:
        mm = torch.stack([coords.min(dim=0).values, coords.max(dim=0).values]).cpu().tolist()
        b = (tuple(int(v) for v in mm[0]), tuple(int(v) for v in mm[1]))
    if cache is not None:
        cache[BOUNDS_KEY] = b
    return b


def set_coord_bounds(cache: Dict, lo: Sequence[int], hi: Sequence[int]) -> None:
    cache[BOUNDS_KEY] = (tuple(int(v) for v in lo), tuple(int(v) for v in hi))


def make_keyspec(bounds, div: Sequence[int], order: Sequence[int], mul: Sequence[int] = (1, 1, 1),
                 pad: int = 0) -> Tuple[_capi.KeySpec, int]:
    """Field layout for lk_pack_keys given coordinate bounds.  `pad` widens each spatial field's
    range by that many cells on both sides (neighbour offsets must stay representable)."""
    lo, hi = bounds
    spec = _capi.KeySpec()
    total = 0
    for a in range(4):
        if a < 3:
            d = int(div[a])
            qlo, qhi = lo[a] // d - pad, hi[a] // d + pad     # python // is floor division
            spec.div[a] = d
            spec.mul[a] = int(mul[a])
        else:
            qlo, qhi = lo[3], hi[3]
        spec.lo[a] = qlo
        spec.bits[a] = max(int(qhi - qlo).bit_length(), 0)
        total += spec.bits[a]
    for f in range(4):
        spec.order[f] = int(order[f])
    if total > 64:
        raise RuntimeError(f'coordinate range needs {total} key bits (> 64)')
    return spec, max(total, 1)


def pack_keys(coords: torch.Tensor, spec: _capi.KeySpec) -> torch.Tensor:
    n = coords.shape[0]
    keys = torch.empty(n, dtype=torch.int64, device=coords.device)
    _capi.check(_capi.lib().lk_pack_keys(_capi.ptr(coords, torch.int32), n, C.byref(spec),
                                         _capi.ptr(keys), _capi.stream()), 'lk_pack_keys')
    return keys


def unpack_keys(keys: torch.Tensor, n: int, spec: _capi.KeySpec,
                d_count: Optional[torch.Tensor] = None) -> torch.Tensor:
    coords = torch.empty(n, 4, dtype=torch.int32, device=keys.device)
    _capi.check(_capi.lib().lk_unpack_keys(_capi.ptr(keys), _capi.ptr(d_count), n, C.byref(spec),
                                           _capi.ptr(coords), _capi.stream()), 'lk_unpack_keys')
    return coords


class SortUnique:
    """Result of lk_sort_unique: capacity-n buffers + the device scalar `num`."""
    __slots__ = ('unique', 'inverse', 'order', 'seg', 'counts', 'num', 'n', 'sorted_rank')


def sort_unique(keys: torch.Tensor, key_bits: int, want_order: bool = False) -> SortUnique:
    n = keys.shape[0]
    dev = keys.device
    r = SortUnique()
    r.n = n
    r.unique = torch.empty(n, dtype=torch.int64, device=dev)
    r.inverse = torch.empty(n, dtype=torch.int32, device=dev)
    r.counts = torch.empty(n, dtype=torch.int32, device=dev)
    r.order = torch.empty(n, dtype=torch.int32, device=dev) if want_order else None
    r.seg = torch.empty(n + 1, dtype=torch.int32, device=dev) if want_order else None
    r.sorted_rank = torch.empty(n, dtype=torch.int32, device=dev) if want_order else None
    r.num = torch.empty(1, dtype=torch.int32, device=dev)
    L = _capi.lib()
    ws_bytes = L.lk_sort_unique_ws_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with _capi.timed('lk_sort_unique', n * 12 * 2 * ((int(key_bits) + 7) // 8)):
        _capi.check(L.lk_sort_unique_ex(_capi.ptr(keys), n, int(key_bits), _capi.ptr(r.unique),
                                        _capi.ptr(r.inverse), _capi.ptr(r.order), _capi.ptr(r.seg),
                                        _capi.ptr(r.counts), _capi.ptr(r.num),
                                        _capi.ptr(r.sorted_rank), _capi.ptr(ws), ws_bytes,
                                        _capi.stream()), 'lk_sort_unique_ex')
    return r


def unique_coords(coords: torch.Tensor, div: Sequence[int], order: Sequence[int],
                  mul: Sequence[int] = (1, 1, 1), cache: Optional[Dict] = None):
    """Device replacement for `torch.unique(q, dim=0)` where q = (coords[:, :3] // div, batch),
    sorted lexicographically by the fields in `order`.  Returns (unique coords [M,4] int32 with
    spatial fields multiplied by `mul`, inverse [N] int32, counts [M] int32)."""
    bounds = coord_bounds(coords, cache)
    spec, bits = make_keyspec(bounds, div, order, mul)
    keys = pack_keys(coords, spec)
    su = sort_unique(keys, bits)
    m = int(su.num.item())          # the one size read-back a shape-returning API needs
    uc = unpack_keys(su.unique, m, spec)
    return uc, su.inverse, su.counts[:m]
