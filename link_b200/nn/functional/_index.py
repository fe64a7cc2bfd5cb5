"""Host-side helpers around the device index primitives (key packing, radix sort/unique).

The only host<->device synchronisation of the whole index path is `coord_bounds` (one 8-int
read-back per input scan, cached in the tensor family's shared `kmaps` dict); callers that
need CUDA-graph capture can pre-seed the cache with `set_coord_bounds`."""
import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import torch

from link_b200 import _capi

BOUNDS_KEY = ('lk', 'bounds')
# LINKB200_CHECK_BOUNDS=1: verify seeded / derived bounds against the tensor on the host (one sync)
import os as _os
CHECK_BOUNDS = _os.environ.get('LINKB200_CHECK_BOUNDS', '0') == '1'


def _tkey(coords: torch.Tensor):
    return ('lk', 'bounds', coords.data_ptr(), coords.shape[0])


def _exact_bounds(coords: torch.Tensor):
    if coords.shape[0] == 0:
        return ((0, 0, 0, 0), (0, 0, 0, 0))
    mm = torch.stack([coords.min(dim=0).values, coords.max(dim=0).values]).cpu().tolist()
    return (tuple(int(v) for v in mm[0]), tuple(int(v) for v in mm[1]))


def coord_bounds(coords: torch.Tensor, cache: Optional[Dict] = None):
    """(lo[4], hi[4]) python ints of an int32 [N,4] coordinate tensor.

    With `cache` (a SparseTensor.kmaps dict) bounds are remembered PER COORDINATE TENSOR (keyed by
    storage address and length).  Levels that our own ops derive (strided convs) get their bounds
    analytically from the parent's (`register_floored_bounds`), so a forward pass pays one 8-int
    read-back for the input scan and none for the derived levels.  The family's seed entry
    (`set_coord_bounds`, or the first tensor measured) is bound to the first tensor that asks for it
    and is never applied to a different, unregistered tensor: derived coordinates may lie outside
    the finest level's bounds (aligned outputs of a k != stride downsample above `hi`, floors below
    `lo` when a block edge is not a multiple of the tensor stride), and a field outside its key
    range would alias silently."""
    if cache is None:
        return _exact_bounds(coords)
    key = _tkey(coords)
    hit = cache.get(key)
    if hit is not None:
        return hit
    seed = cache.get(BOUNDS_KEY)
    if seed is not None and len(seed) == 2:          # pre-seeded, not bound yet: this is the input scan
        b = (seed[0], seed[1])
        cache[BOUNDS_KEY] = (seed[0], seed[1], key)
    else:
        b = _exact_bounds(coords)
        if seed is None:
            cache[BOUNDS_KEY] = (b[0], b[1], key)
    if CHECK_BOUNDS:
        ex = _exact_bounds(coords)
        if any(ex[0][a] < b[0][a] or ex[1][a] > b[1][a] for a in range(4)):
            raise RuntimeError(f'coordinate bounds {b} do not contain the tensor\'s range {ex}')
    cache[key] = b
    return b


def set_coord_bounds(cache: Dict, lo: Sequence[int], hi: Sequence[int]) -> None:
    """Pre-seed the bounds of the input scan (a dataset property the loader knows): the first
    coordinate tensor of the family that needs bounds takes them without a device read-back."""
    cache[BOUNDS_KEY] = (tuple(int(v) for v in lo), tuple(int(v) for v in hi))


def register_floored_bounds(cache: Optional[Dict], child: torch.Tensor, parent_bounds, step: Sequence[int]) -> None:
    """Bounds of coordinates derived as floor(c / step) * step from a tensor with `parent_bounds`."""
    if cache is None:
        return
    lo, hi = parent_bounds
    st = [int(v) for v in step] + [1]
    cache[_tkey(child)] = (tuple((lo[a] // st[a]) * st[a] for a in range(4)),
                           tuple((hi[a] // st[a]) * st[a] for a in range(4)))


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
