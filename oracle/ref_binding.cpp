// pybind11 binding (ours) for the reference's boxes_iou_bev_cpu; the implementation is compiled from
// /root/reference/utils/detzero_utils/ops/iou3d_nms/src/iou3d_cpu.cpp in place.
#include <torch/extension.h>
int boxes_iou_bev_cpu(at::Tensor boxes_a_tensor, at::Tensor boxes_b_tensor, at::Tensor ans_iou_tensor);
PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) { m.def("boxes_iou_bev_cpu", &boxes_iou_bev_cpu, "reference rotated BEV IoU (CPU)"); }
