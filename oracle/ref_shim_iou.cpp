// oracle/ref_shim_iou.cpp -- TEST INFRASTRUCTURE.  A C-ABI window onto the reference's own
// rotated-IoU header, compiled IN PLACE from /root/reference (never copied into this repo):
//   -I/root/reference/utils/box_iou_rotated/src      -> 5-float boxes   (REF_ML undefined)
//   -I/root/reference/utils/ml_nms_rotated/src       -> 6-float boxes   (-DREF_ML)
// Built twice by oracle/build_ref.py: with g++ (the reference's CPU semantics: std::sort hull
// ordering) and with nvcc as host code (__CUDACC__ defined -> the exchange-sort hull ordering
// that the reference's CUDA kernels execute).  Outputs go to oracle/_ref/ only.
#include <iostream>
#include <cstdint>
#include "box_iou_rotated_utils.h"

#ifndef REF_SYM
#error "define REF_SYM(name) to give the exported symbols a unique prefix"
#endif

#ifdef REF_ML
static const int kStride = 6;
#else
static const int kStride = 5;
#endif

extern "C" {

float REF_SYM(single_iou)(const float* b1, const float* b2) {
  return single_box_iou_rotated<float>(b1, b2);
}

// out[i*m + j] = iou(boxes1[i], boxes2[j]); boxes are kStride floats apart.
void REF_SYM(pairwise)(const float* boxes1, int64_t n, const float* boxes2, int64_t m, float* out) {
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = 0; j < m; ++j)
      out[i * m + j] = single_box_iou_rotated<float>(boxes1 + kStride * i, boxes2 + kStride * j);
}

}  // extern "C"
