"""points_to_voxel on the GPU, same signature and return values as the reference
(det3d/ops/point_cloud/point_cloud_ops.py:112-183), which runs a single-threaded numba loop inside
the data loader.  Results are identical: voxels in order of first appearance, at most `max_voxels`,
the first `max_points` points of each voxel in point order, coordinates (z, y, x) when
`reverse_index` (the only mode the reference's VoxelGenerator uses, core/input/voxel_generator.py:23)."""
import ctypes as C

import numpy as np
import torch

from link_b200 import _capi

__all__ = ['points_to_voxel']


def points_to_voxel(points: torch.Tensor, voxel_size, coors_range, max_points: int = 35,
                    reverse_index: bool = True, max_voxels: int = 20000):
    """points [N, ndim >= 3] float32 CUDA tensor -> (voxels [M, max_points, ndim], coordinates [M, 3]
    int32, num_points_per_voxel [M] int32), all on the device."""
    if not points.is_cuda:
        raise RuntimeError('link_b200.ops.points_to_voxel needs a CUDA tensor (there is no CPU fallback)')
    points = points.contiguous().float()
    n, ndim = points.shape
    vs = np.ascontiguousarray(np.asarray(voxel_size, dtype=np.float32))
    cr = np.ascontiguousarray(np.asarray(coors_range, dtype=np.float32))
    dev = points.device
    voxels = torch.empty(max_voxels, max_points, ndim, dtype=torch.float32, device=dev)
    coors = torch.empty(max_voxels, 3, dtype=torch.int32, device=dev)
    num = torch.empty(max_voxels, dtype=torch.int32, device=dev)
    vnum = torch.empty(1, dtype=torch.int32, device=dev)
    L = _capi.lib()
    ws_bytes = L.lk_points_to_voxel_ws_bytes(n)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    _capi.check(L.lk_points_to_voxel(_capi.ptr(points), n, ndim, vs.ctypes.data_as(C.c_void_p),
                                     cr.ctypes.data_as(C.c_void_p), int(max_points), int(max_voxels),
                                     _capi.ptr(voxels), _capi.ptr(coors), _capi.ptr(num), _capi.ptr(vnum),
                                     _capi.ptr(ws), ws_bytes, _capi.stream()), 'lk_points_to_voxel')
    m = int(vnum.item())            # the one size read-back a shape-returning API needs
    coors = coors[:m]
    if not reverse_index:
        coors = coors.flip(1).contiguous()
    return voxels[:m], coors, num[:m]
