// ref_shim_poly.cpp -- TEST INFRASTRUCTURE: a C entry point around the reference's own iou_poly
// (DOTA_devkit/polyiou/csrc/polyiou.cpp:110-126), which oracle/build_oracle.py compiles IN PLACE from
// /root/reference together with this file into oracle/_ref/libref_polyiou.so.  Nothing of the reference is copied.
#include <vector>

double iou_poly(std::vector<double> p, std::vector<double> q);

extern "C" __attribute__((visibility("default"))) double ref_iou_poly(const double* p, const double* q) {
  return iou_poly(std::vector<double>(p, p + 8), std::vector<double>(q, q + 8));
}

extern "C" __attribute__((visibility("default"))) void ref_iou_poly_pairs(const double* p, const double* q, long long n,
                                                                         double* out) {
  for (long long i = 0; i < n; ++i) out[i] = ref_iou_poly(p + 8 * i, q + 8 * i);
}
