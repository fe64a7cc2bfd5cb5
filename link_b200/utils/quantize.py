"""Host-side voxelisation front-end with the reference's numpy semantics
(torchsparse/utils/quantize.py:9-46): runs in data-loader workers, before tensors reach
the GPU.  The device-side equivalent is link_b200.nn.functional.sparse_quantize_cuda."""
from itertools import repeat
from typing import List, Tuple, Union

import numpy as np

__all__ = ['sparse_quantize', 'ravel_hash']


def ravel_hash(x: np.ndarray) -> np.ndarray:
    assert x.ndim == 2, x.shape
    x -= np.min(x, axis=0)          # in place, like the reference (callers rely on the shift)
    x = x.astype(np.uint64, copy=False)
    xmax = np.max(x, axis=0).astype(np.uint64) + 1
    h = np.zeros(x.shape[0], dtype=np.uint64)
    for k in range(x.shape[1] - 1):
        h += x[:, k]
        h *= xmax[k + 1]
    h += x[:, -1]
    return h


def sparse_quantize(coords, voxel_size: Union[float, Tuple[float, ...]] = 1, *,
                    return_index: bool = False, return_inverse: bool = False) -> List[np.ndarray]:
    if isinstance(voxel_size, (float, int)):
        voxel_size = tuple(repeat(voxel_size, 3))
    assert isinstance(voxel_size, tuple) and len(voxel_size) == 3
    coords = np.floor(coords / np.array(voxel_size)).astype(np.int32)
    _, indices, inverse_indices = np.unique(ravel_hash(coords), return_index=True,
                                            return_inverse=True)
    coords = coords[indices]
    outputs = [coords]
    if return_index:
        outputs += [indices]
    if return_inverse:
        outputs += [inverse_indices.reshape(-1)]
    return outputs[0] if len(outputs) == 1 else outputs
