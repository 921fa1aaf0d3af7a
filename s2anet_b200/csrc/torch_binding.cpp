// torch_binding.cpp -- the thin torch extension over the C ABI of libs2a_b200.so (the reference's
// boundary is six pybind11 torch-extension modules; this module exports the same function names with the same positional
// signatures for the latency-sensitive ops, so `dropin.install()` can register it under the reference's module names).
//
// Every function: checks its arguments the way the reference hosts do (RuntimeError from TORCH_CHECK, like the
// reference's AT_ASSERTM / TORCH_CHECK), allocates the result with ATen on the inputs' device, and calls ONE C-ABI entry
// on the current CUDA stream.  No kernel code lives here.
//
// reference: utils/box_iou_rotated/src/box_iou_rotated.h:22-42 + box_iou_rotated_cuda.cu:65-101,
//            utils/nms_rotated/src/nms_rotated.h:21-40 + nms_rotated_cuda.cu:72-132,
//            utils/ml_nms_rotated/src/nms_rotated.h:23-43 + nms_rotated_cuda.cu:74-137,
//            models/orn/src/vision.cpp:7-12 + ActiveRotatingFilter.h:11-43 (arf_forward / arf_backward)
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "s2a_b200.h"

namespace {

void check(int rc, const char* what) {
  TORCH_CHECK(rc == S2A_OK, "s2anet_b200: ", what, " failed (status ", rc, "): ", s2a_last_error());
}

void* current_stream() { return (void*)at::cuda::getCurrentCUDAStream().stream(); }

int dtype_code(const at::Tensor& t, const char* what) {
  switch (t.scalar_type()) {
    case at::kFloat: return S2A_F32;
    case at::kBFloat16: return S2A_BF16;
    case at::kHalf: return S2A_F16;
    default: TORCH_CHECK(false, what, ": unsupported dtype ", t.scalar_type(), " (float32, float16, bfloat16)");
  }
  return -1;
}

// utils/box_iou_rotated/src/box_iou_rotated.h:22-42.  Strided / non-fp32 inputs are accepted (the reference CUDA op
// reads raw pointers and silently mis-reads them, models/utils.py:51-56).
at::Tensor box_iou_rotated(const at::Tensor& boxes1, const at::Tensor& boxes2) {
  TORCH_CHECK(boxes1.is_cuda() && boxes2.is_cuda(), "box_iou_rotated: boxes must be CUDA tensors (this library has no CPU path)");
  TORCH_CHECK(boxes1.device() == boxes2.device(), "box_iou_rotated: boxes1 and boxes2 are on different devices");
  const int64_t n = boxes1.size(0), m = boxes2.size(0);
  c10::cuda::CUDAGuard guard(boxes1.device());
  at::Tensor out = at::empty({n, m}, boxes1.options().dtype(at::kFloat));
  if (n == 0 || m == 0) return out;
  TORCH_CHECK(boxes1.dim() == 2 && boxes2.dim() == 2 && boxes1.size(1) == 5 && boxes2.size(1) == 5,
              "box_iou_rotated: boxes must be [N,5] and [M,5]");
  const at::Tensor b1 = boxes1.to(at::kFloat).contiguous(), b2 = boxes2.to(at::kFloat).contiguous();
  check(s2a_box_iou_rotated(b1.data_ptr<float>(), n, b2.data_ptr<float>(), m, 1, out.data_ptr<float>(), m, 0, n,
                            S2A_IOU_DEFAULT, current_stream()),
        "box_iou_rotated");
  return out;
}

// nms_rotated (labels undefined) / ml_nms_rotated: keep indices into the input, descending score, int64, on the device.
// The only host synchronisation is the 4-byte read of the result length.
at::Tensor nms_impl(const at::Tensor& dets_in, const at::Tensor& scores_in, const c10::optional<at::Tensor>& labels_in,
                    double iou_threshold, const char* what) {
  TORCH_CHECK(dets_in.is_cuda() && scores_in.is_cuda(), what, ": dets and scores must be CUDA tensors");
  c10::cuda::CUDAGuard guard(dets_in.device());
  const int64_t n = dets_in.size(0);
  if (n == 0) return at::empty({0}, dets_in.options().dtype(at::kLong));
  TORCH_CHECK(dets_in.dim() == 2 && dets_in.size(1) == 5, what, ": dets must be [N,5]");
  TORCH_CHECK(scores_in.dim() == 1 && scores_in.size(0) == n, what, ": scores must be [N]");
  at::Tensor dets = dets_in, scores = scores_in, labels;
  if (dets.scalar_type() != at::kFloat || dets.stride(1) != 1) dets = dets.to(at::kFloat).contiguous();
  if (scores.scalar_type() != at::kFloat) scores = scores.to(at::kFloat);      // fp16 scores under half validation
  const float* lp = nullptr;
  if (labels_in.has_value()) {
    labels = *labels_in;
    TORCH_CHECK(labels.is_cuda() && labels.dim() == 1 && labels.size(0) == n, what, ": labels must be a CUDA tensor [N]");
    if (labels.scalar_type() != at::kFloat || labels.stride(0) != 1) labels = labels.to(at::kFloat).contiguous();
    lp = labels.data_ptr<float>();
  }
  const size_t ws_bytes = s2a_nms_rotated_workspace_bytes(n);
  at::Tensor ws = at::empty({(int64_t)ws_bytes}, dets.options().dtype(at::kByte));
  at::Tensor keep = at::empty({n}, dets.options().dtype(at::kLong));
  at::Tensor num = at::empty({1}, dets.options().dtype(at::kInt));
  check(s2a_nms_rotated(dets.data_ptr<float>(), dets.stride(0), scores.data_ptr<float>(), scores.stride(0), lp, n,
                        (float)iou_threshold, keep.data_ptr<int64_t>(), num.data_ptr<int32_t>(), ws.data_ptr(), ws_bytes,
                        current_stream()),
        what);
  return keep.narrow(0, 0, (int64_t)num.item<int32_t>());
}

at::Tensor nms_rotated(const at::Tensor& dets, const at::Tensor& scores, double iou_threshold) {
  return nms_impl(dets, scores, c10::nullopt, iou_threshold, "nms_rotated");
}

at::Tensor ml_nms_rotated(const at::Tensor& dets, const at::Tensor& scores, const at::Tensor& labels, double iou_threshold) {
  return nms_impl(dets, scores, labels, iou_threshold, "ml_nms_rotated");
}

// models/orn/src/ActiveRotatingFilter.h:11-21: weight [O, I, nOri, kH, kW], indices uint8 [nOri, kH, kW, nRot] ->
// [O*nRot, I*nOri, kH, kW]
at::Tensor arf_forward(const at::Tensor& weight, const at::Tensor& indices) {
  TORCH_CHECK(weight.is_cuda() && indices.is_cuda(), "arf_forward: tensors must be CUDA tensors");
  TORCH_CHECK(weight.dim() == 5, "arf_forward: only supports a batch of ARFs (weight must be 5-D)");
  TORCH_CHECK(indices.dim() == 4 && indices.scalar_type() == at::kByte, "arf_forward: indices must be uint8 [nOri,kH,kW,nRot]");
  const int64_t O = weight.size(0), I = weight.size(1), nOri = weight.size(2), kH = weight.size(3), kW = weight.size(4);
  const int64_t nRot = indices.size(3);
  TORCH_CHECK(indices.size(0) == nOri && indices.size(1) == kH && indices.size(2) == kW, "arf_forward: indices do not match the weight");
  c10::cuda::CUDAGuard guard(weight.device());
  const at::Tensor w = weight.contiguous(), idx = indices.contiguous();
  at::Tensor out = at::empty({O * nRot, I * nOri, kH, kW}, w.options());
  if (out.numel() == 0) return out;
  check(s2a_arf_forward(w.data_ptr(), idx.data_ptr<uint8_t>(), out.data_ptr(), (int)O, (int)I, (int)nOri, (int)kH, (int)kW,
                        (int)nRot, dtype_code(w, "arf_forward"), current_stream()),
        "arf_forward");
  return out;
}

// models/orn/src/ActiveRotatingFilter.h:23-43: gradOutput [O*nRot, I*nOri, kH, kW] -> gradWeight [O, I, nOri, kH, kW]
at::Tensor arf_backward(const at::Tensor& indices, const at::Tensor& grad_output) {
  TORCH_CHECK(grad_output.is_cuda() && indices.is_cuda(), "arf_backward: tensors must be CUDA tensors");
  TORCH_CHECK(indices.dim() == 4 && indices.scalar_type() == at::kByte, "arf_backward: indices must be uint8 [nOri,kH,kW,nRot]");
  TORCH_CHECK(grad_output.dim() == 4, "arf_backward: gradOutput must be 4-D");
  const int64_t nOri = indices.size(0), kH = indices.size(1), kW = indices.size(2), nRot = indices.size(3);
  TORCH_CHECK(nRot > 0 && nOri > 0 && grad_output.size(0) % nRot == 0 && grad_output.size(1) % nOri == 0 &&
                  grad_output.size(2) == kH && grad_output.size(3) == kW,
              "arf_backward: gradOutput does not match the indices");
  const int64_t O = grad_output.size(0) / nRot, I = grad_output.size(1) / nOri;
  c10::cuda::CUDAGuard guard(grad_output.device());
  const at::Tensor g = grad_output.contiguous(), idx = indices.contiguous();
  at::Tensor gw = at::empty({O, I, nOri, kH, kW}, g.options());
  if (gw.numel() == 0) return gw;
  check(s2a_arf_backward(g.data_ptr(), idx.data_ptr<uint8_t>(), gw.data_ptr(), (int)O, (int)I, (int)nOri, (int)kH, (int)kW,
                         (int)nRot, dtype_code(g, "arf_backward"), current_stream()),
        "arf_backward");
  return gw;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "s2anet_b200: torch-extension binding of libs2a_b200.so (box_iou_rotated, nms_rotated, ml_nms_rotated, arf_*)";
  m.def("box_iou_rotated", &box_iou_rotated, "IoU for rotated boxes", py::arg("boxes1"), py::arg("boxes2"));
  m.def("nms_rotated", &nms_rotated, "NMS for rotated boxes", py::arg("dets"), py::arg("scores"), py::arg("iou_threshold"));
  m.def("ml_nms_rotated", &ml_nms_rotated, "multi-label NMS for rotated boxes", py::arg("dets"), py::arg("scores"),
        py::arg("labels"), py::arg("iou_threshold"));
  m.def("arf_forward", &arf_forward, "ARF forward", py::arg("weight"), py::arg("indices"));
  m.def("arf_backward", &arf_backward, "ARF backward", py::arg("indices"), py::arg("gradOutput"));
  m.def("abi_version", []() { return s2a_version(); });
}
