"""Host-side voxelisation front-end (data-loader workers, before tensors reach the GPU).

Contract = the reference's `sparse_quantize` (torchsparse/utils/quantize.py:24-46): points are
floored to a voxel grid, duplicates collapse to the FIRST point of each voxel, voxels come out in
ascending x-major (x, then y, then z) order, and -- because the reference's `ravel_hash` shifts its
argument in place (quantize.py:12) -- the returned coordinates are relative to the lower corner of
the scan.  Written here as one stable sort of a mixed-radix key instead of `np.unique` on a hash
(same outputs: pinned by tests/golden/voxelize.npz).  The device-side counterparts are
`elk.initial_voxelize` and `ops.point_cloud_ops.points_to_voxel`."""
from typing import List, Sequence, Union

import numpy as np

__all__ = ['sparse_quantize', 'voxel_keys']


def voxel_keys(grid: np.ndarray) -> np.ndarray:
    """Mixed-radix int64 key of non-negative integer rows: the first column is the most significant
    digit, so ascending keys = lexicographic order of the rows."""
    if grid.ndim != 2:
        raise ValueError(f'expected [N, D] integer coordinates, got shape {grid.shape}')
    g = grid.astype(np.int64, copy=False)
    if g.size and g.min() < 0:
        raise ValueError('voxel_keys needs non-negative coordinates (shift by the minimum first)')
    extent = g.max(axis=0) + 1 if g.shape[0] else np.ones(g.shape[1], np.int64)
    radix = np.ones(g.shape[1], dtype=np.int64)
    for d in range(g.shape[1] - 2, -1, -1):          # radix[d] = product of the extents to the right
        radix[d] = radix[d + 1] * extent[d + 1]
    return g @ radix


def sparse_quantize(coords, voxel_size: Union[float, Sequence[float]] = 1, *,
                    return_index: bool = False, return_inverse: bool = False) -> List[np.ndarray]:
    size = np.asarray(voxel_size, dtype=np.float64)
    if size.ndim == 0:
        size = np.full(3, float(size))
    if size.shape != (3,):
        raise AssertionError('voxel_size must be a number or a 3-tuple')
    grid = np.floor(np.asarray(coords) / size).astype(np.int32)
    n = grid.shape[0]
    if n == 0:
        empty = np.zeros(0, dtype=np.int64)
        out = [grid] + ([empty] if return_index else []) + ([empty] if return_inverse else [])
        return out[0] if len(out) == 1 else out
    grid -= grid.min(axis=0)
    key = voxel_keys(grid)
    order = np.argsort(key, kind='stable')            # stable: the first point of a voxel stays first
    sorted_key = key[order]
    head = np.ones(n, dtype=bool)
    head[1:] = sorted_key[1:] != sorted_key[:-1]
    first = order[head]                               # row of the first point of every voxel
    out = [grid[first]]
    if return_index:
        out.append(first)
    if return_inverse:
        inverse = np.empty(n, dtype=np.int64)
        inverse[order] = np.cumsum(head) - 1          # voxel number of every point
        out.append(inverse)
    return out[0] if len(out) == 1 else out
