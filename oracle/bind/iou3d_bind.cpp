// Python binding for the reference's CPU rotated-BEV-IoU (test infrastructure only).
// boxes_iou_bev_cpu is defined in /root/reference/detection/det3d/ops/iou3d_nms/src/iou3d_cpu.cpp,
// which is compiled in place next to this file by oracle/build_ref.py::build_iou3d (the reference
// binds it in iou3d_nms_api.cpp:13, together with CUDA entry points that cannot be built here).
#include <torch/extension.h>

int boxes_iou_bev_cpu(at::Tensor boxes_a_tensor, at::Tensor boxes_b_tensor, at::Tensor ans_iou_tensor);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("boxes_iou_bev_cpu", &boxes_iou_bev_cpu, "rotated BEV IoU of [N,7] x [M,7] boxes (x, y, z, dx, dy, dz, heading)");
}
