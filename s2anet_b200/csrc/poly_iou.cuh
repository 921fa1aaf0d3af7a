// poly_iou.cuh -- IoU of two quadrilaterals in fp64, the arithmetic of the DOTA result-merging NMS.
//
// Replaces (reference): DOTA_devkit/polyiou/csrc/polyiou.cpp:9-126 (sig, cross, area, lineCross, polygon_cut,
// the two intersectArea overloads, iou_poly), called per pair from py_cpu_nms_poly_fast
// (DOTA_devkit/ResultMerge_multi_process.py:62-123).
//
// The intersection of two polygons is the sum over all (edge of P, edge of Q) pairs of the SIGNED intersection
// area of the triangles (origin, edge); each triangle pair is a triangle clipped by three half planes.  Every
// floating-point operation is issued individually rounded, in the reference's order, so the device result is the
// reference's double bit for bit (the reference is built without FMA contraction for the x86-64 baseline).
// All functions are __host__ __device__: tests/host_harness.cu runs the same code on the CPU.
//
// One corner of the reference is undefined and stays so: polygon_cut appends `pp[m++]` even when lineCross finds
// no crossing (|s2 - s1| <= 1e-8 with different signs of s1, s2) and then reads an uninitialised point; here that
// slot keeps whatever the previous cut of the same triangle pair left there (zero at first).
#pragma once
#include <math.h>

namespace s2a {

#define PI_HD __host__ __device__ __forceinline__

#ifdef __CUDA_ARCH__
#define PI_MUL(a, b) __dmul_rn((a), (b))
#define PI_ADD(a, b) __dadd_rn((a), (b))
#define PI_SUB(a, b) __dsub_rn((a), (b))
#define PI_DIV(a, b) __ddiv_rn((a), (b))
#else
#define PI_MUL(a, b) ((a) * (b))
#define PI_ADD(a, b) ((a) + (b))
#define PI_SUB(a, b) ((a) - (b))
#define PI_DIV(a, b) ((a) / (b))
#endif

struct PPt { double x, y; };

PI_HD int pi_sig(double d) { return (d > 1e-8) - (d < -1e-8); }                    // polyiou.cpp:9-11
PI_HD bool pi_same(const PPt& a, const PPt& b) {                                   // :15-17
  return pi_sig(PI_SUB(a.x, b.x)) == 0 && pi_sig(PI_SUB(a.y, b.y)) == 0;
}
PI_HD double pi_cross(const PPt& o, const PPt& a, const PPt& b) {                   // :19-21
  return PI_SUB(PI_MUL(PI_SUB(a.x, o.x), PI_SUB(b.y, o.y)), PI_MUL(PI_SUB(b.x, o.x), PI_SUB(a.y, o.y)));
}
// :22-29 (writes ps[n] = ps[0] like the reference)
PI_HD double pi_area(PPt* ps, int n) {
  ps[n] = ps[0];
  double res = 0.0;
  for (int i = 0; i < n; ++i) res = PI_ADD(res, PI_SUB(PI_MUL(ps[i].x, ps[i + 1].y), PI_MUL(ps[i].y, ps[i + 1].x)));
  return PI_DIV(res, 2.0);
}
// :30-39; the caller guarantees sig(s1) != sig(s2), so the "2" (collinear) outcome cannot occur there
PI_HD int pi_line_cross(const PPt& a, const PPt& b, const PPt& c, const PPt& d, PPt& p) {
  const double s1 = pi_cross(a, b, c), s2 = pi_cross(a, b, d);
  if (pi_sig(s1) == 0 && pi_sig(s2) == 0) return 2;
  const double den = PI_SUB(s2, s1);
  if (pi_sig(den) == 0) return 0;
  p.x = PI_DIV(PI_SUB(PI_MUL(c.x, s2), PI_MUL(d.x, s1)), den);
  p.y = PI_DIV(PI_SUB(PI_MUL(c.y, s2), PI_MUL(d.y, s1)), den);
  return 1;
}
// :58-71 -- keep the part of polygon p (n vertices, room for n + 1) on the left of a->b, in place
PI_HD void pi_polygon_cut(PPt* p, int& n, const PPt& a, const PPt& b, PPt* pp) {
  int m = 0;
  p[n] = p[0];
  for (int i = 0; i < n; ++i) {
    const int si = pi_sig(pi_cross(a, b, p[i]));
    if (si > 0) pp[m++] = p[i];
    if (si != pi_sig(pi_cross(a, b, p[i + 1]))) pi_line_cross(a, b, p[i], p[i + 1], pp[m++]);
  }
  n = 0;
  for (int i = 0; i < m; ++i)
    if (!i || !pi_same(pp[i], pp[i - 1])) p[n++] = pp[i];
  while (n > 1 && pi_same(p[n - 1], p[0])) --n;
}
// :74-90 -- signed intersection area of the triangles (o, a, b) and (o, c, d), o = origin
PI_HD double pi_tri_intersect(PPt a, PPt b, PPt c, PPt d) {
  const PPt o = {0.0, 0.0};
  const int s1 = pi_sig(pi_cross(o, a, b)), s2 = pi_sig(pi_cross(o, c, d));
  if (s1 == 0 || s2 == 0) return 0.0;
  if (s1 == -1) { const PPt t = a; a = b; b = t; }
  if (s2 == -1) { const PPt t = c; c = d; d = t; }
  PPt p[10], pp[20];
  for (int i = 0; i < 20; ++i) pp[i] = o;
  p[0] = o; p[1] = a; p[2] = b;
  int n = 3;
  pi_polygon_cut(p, n, o, c, pp);
  pi_polygon_cut(p, n, c, d, pp);
  pi_polygon_cut(p, n, d, o, pp);
  double res = fabs(pi_area(p, n));
  if (s1 * s2 == -1) res = -res;
  return res;
}
// :92-105 + :110-126 -- p, q: 8 doubles each (x0, y0, ..., x3, y3)
PI_HD double poly_iou(const double* p, const double* q) {
  PPt ps1[5], ps2[5];
  for (int i = 0; i < 4; ++i) {
    ps1[i].x = p[2 * i]; ps1[i].y = p[2 * i + 1];
    ps2[i].x = q[2 * i]; ps2[i].y = q[2 * i + 1];
  }
  if (pi_area(ps1, 4) < 0.0) { PPt t = ps1[0]; ps1[0] = ps1[3]; ps1[3] = t; t = ps1[1]; ps1[1] = ps1[2]; ps1[2] = t; }
  if (pi_area(ps2, 4) < 0.0) { PPt t = ps2[0]; ps2[0] = ps2[3]; ps2[3] = t; t = ps2[1]; ps2[1] = ps2[2]; ps2[2] = t; }
  ps1[4] = ps1[0];
  ps2[4] = ps2[0];
  double inter = 0.0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) inter = PI_ADD(inter, pi_tri_intersect(ps1[i], ps1[i + 1], ps2[j], ps2[j + 1]));
  const double uni = PI_SUB(PI_ADD(fabs(pi_area(ps1, 4)), fabs(pi_area(ps2, 4))), inter);
  return PI_DIV(inter, uni);
}

// py_cpu_nms_poly_fast's pair value (ResultMerge_multi_process.py:64-69, 89-104): the axis-aligned boxes of the two
// polygons first -- area (x2 - x1 + 1)(y2 - y1 + 1), overlap max(0, .) without the +1 -- and the polygon IoU only
// where that overlap ratio is > 0.  hb = (x1, y1, x2, y2, area).
PI_HD double poly_nms_pair(const double* pi, const double* hbi, const double* pj, const double* hbj) {
  const double xx1 = fmax(hbi[0], hbj[0]), yy1 = fmax(hbi[1], hbj[1]);
  const double xx2 = fmin(hbi[2], hbj[2]), yy2 = fmin(hbi[3], hbj[3]);
  const double w = fmax(0.0, PI_SUB(xx2, xx1)), h = fmax(0.0, PI_SUB(yy2, yy1));
  const double hbb_inter = PI_MUL(w, h);
  const double hbb_ovr = PI_DIV(hbb_inter, PI_SUB(PI_ADD(hbi[4], hbj[4]), hbb_inter));
  if (!(hbb_ovr > 0.0)) return hbb_ovr;
  return poly_iou(pi, pj);
}

}  // namespace s2a
