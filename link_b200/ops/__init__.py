"""Device versions of the detection pipeline's point-cloud ops (det3d/ops/point_cloud)."""
from .point_cloud_ops import points_to_voxel  # noqa: F401
