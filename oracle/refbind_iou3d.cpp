// TEST INFRASTRUCTURE (oracle/_ref).  Python binding for the reference's own CPU implementation of the rotated BEV IoU,
// det3d/ops/iou3d_nms/src/iou3d_cpu.cpp:232 (boxes_iou_bev_cpu), compiled from the source where it lies under the
// reference tree by oracle/build_ref.py.  Nothing of the reference is copied here: this file only declares the symbol.
#include <torch/extension.h>

int boxes_iou_bev_cpu(at::Tensor boxes_a_tensor, at::Tensor boxes_b_tensor, at::Tensor ans_iou_tensor);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
    m.def("boxes_iou_bev_cpu", &boxes_iou_bev_cpu, "rotated BEV IoU of (N,7) x (M,7) [x y z dx dy dz heading] boxes -> (N,M), CPU float32");
}
