"""Deterministic synthetic inputs shaped like the reference's datasets (no dataset files, no
network): a spinning-LiDAR simulator for SemanticKITTI-shaped scans and the
uniform random cloud of BASELINE config 1 (G1).  numpy only; used by bench.py, smoke() and tests."""
from typing import Tuple

import numpy as np

from link_b200.utils.quantize import sparse_quantize

__all__ = ['lidar_scan', 'kitti_like_voxels', 'random_voxels']


def lidar_scan(seed: int = 0, beams: int = 64, azimuths: int = 2048, n_boxes: int = 40,
               max_range: float = 50.0) -> np.ndarray:
    """Returns [P,4] float32 (x, y, z, intensity): rays of a `beams` x `azimuths` spinning LiDAR
    at 1.73 m over a ground plane with `n_boxes` axis-aligned boxes, 2 cm range noise."""
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(-24.8, 2.0, beams))
    azim = np.linspace(-np.pi, np.pi, azimuths, endpoint=False)
    el, az = np.meshgrid(elev, azim, indexing='ij')
    d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], -1).reshape(-1, 3)
    origin = np.array([0.0, 0.0, 1.73])
    t = np.full(d.shape[0], np.inf)
    down = d[:, 2] < -1e-6
    t[down] = -origin[2] / d[down, 2]                      # ground plane z = 0
    ctr = np.concatenate([rng.uniform(-max_range, max_range, (n_boxes, 2)),
                          np.zeros((n_boxes, 1))], 1)
    size = rng.uniform(1.0, 10.0, (n_boxes, 3))
    ctr[:, 2] = size[:, 2] / 2
    for c, s in zip(ctr, size):                            # slab test per box
        lo, hi = c - s / 2, c + s / 2
        with np.errstate(divide='ignore', invalid='ignore'):
            t1, t2 = (lo - origin) / d, (hi - origin) / d
        tn = np.nanmax(np.minimum(t1, t2), axis=1)
        tf = np.nanmin(np.maximum(t1, t2), axis=1)
        hit = (tn <= tf) & (tn > 0.5)
        t = np.where(hit & (tn < t), tn, t)
    keep = np.isfinite(t) & (t < max_range * 1.6)
    t = t[keep] + rng.normal(0.0, 0.02, keep.sum())
    pts = origin + d[keep] * t[:, None]
    inten = rng.uniform(0.0, 1.0, (pts.shape[0], 1))
    return np.concatenate([pts, inten], 1).astype(np.float32)


def kitti_like_voxels(n_target: int = 120_000, seed: int = 0, voxel_size: float = 0.05,
                      tol: float = 0.01) -> Tuple[np.ndarray, np.ndarray]:
    """Voxelised synthetic scan with N = n_target (+-tol) active voxels, built the way the
    reference's SemanticKITTI loader does (round(xyz / 0.05) minus the per-axis minimum, then
    sparse_quantize; semantic_kitti.py:219-225).  Returns (coords int32 [N,3] in the loader's
    x-major ravel order, feats float32 [N,4] = (x, y, z, intensity))."""
    az = 2048
    pts = None
    for _ in range(12):                                    # ray density search (deterministic)
        pts = lidar_scan(seed, azimuths=az)
        pc = np.round(pts[:, :3] / voxel_size).astype(np.int32)
        pc -= pc.min(0, keepdims=True)
        _, inds = sparse_quantize(pc.copy(), 1, return_index=True)
        n = len(inds)
        if abs(n - n_target) <= tol * n_target:
            break
        az = max(64, int(round(az * (n_target / n) ** 1.15)))
    pc = np.round(pts[:, :3] / voxel_size).astype(np.int32)
    pc -= pc.min(0, keepdims=True)
    _, inds = sparse_quantize(pc.copy(), 1, return_index=True)
    return np.ascontiguousarray(pc[inds]), np.ascontiguousarray(pts[inds])


def random_voxels(n: int = 8000, extent: int = 64, seed: int = 0, batch: int = 1) -> np.ndarray:
    """BASELINE config 1 (G1): n distinct voxels drawn uniformly in [0, extent)^3, random row
    order, int32 [n*batch, 4] with the batch index in column 3."""
    rng = np.random.default_rng(seed)
    out = []
    for b in range(batch):
        c = np.unique(rng.integers(0, extent, size=(2 * n, 3)), axis=0)
        c = c[rng.permutation(len(c))[:n]]
        out.append(np.concatenate([c, np.full((len(c), 1), b)], 1))
    return np.concatenate(out, 0).astype(np.int32)
