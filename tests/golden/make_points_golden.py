"""Golden fixture for points_to_voxel: runs the reference's own numba implementation
(/root/reference/detection/det3d/ops/point_cloud/point_cloud_ops.py, loaded by path, unmodified) on a
seeded synthetic cloud and stores inputs' seed + outputs.  Build container only."""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = '/root/reference/detection/det3d/ops/point_cloud/point_cloud_ops.py'


def cloud(seed=0, n=20000):
    rng = np.random.default_rng(seed)
    pts = np.concatenate([rng.uniform(-60, 60, (n, 2)), rng.uniform(-6, 4, (n, 1)),
                          rng.uniform(0, 1, (n, 2))], 1).astype(np.float32)
    pts[: n // 3, :2] = rng.normal(0, 4, (n // 3, 2)).astype(np.float32)   # dense centre: many points per voxel
    return pts


def main():
    spec = importlib.util.spec_from_file_location('ref_pco', REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    for tag, (mp, mv) in {'a': (5, 3000), 'b': (10, 120000)}.items():
        pts = cloud(seed=7)
        v, c, n = ref.points_to_voxel(pts, [0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3], mp, True, mv)
        out[f'{tag}_voxels'] = v if tag == 'a' else v[:64]
        out[f'{tag}_voxel_sum'] = v.astype(np.float64).sum(axis=(1, 2))
        out[f'{tag}_coors'], out[f'{tag}_num'] = c, n
        out[f'{tag}_cfg'] = np.array([mp, mv])
    np.savez_compressed(os.path.join(HERE, 'points.npz'), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
