// deform_bwd.cu -- backward pieces of the deformable convolution, fp32, NCHW.
//
// Replaces (reference): deformable_im2col (deform_conv_cuda_kernel.cu:189-275), deformable_col2im (:278-351) and
// deformable_col2im_coord (:353-435), as used by deform_conv_backward_input_cuda / _parameters_cuda
// (models/dcn/src/deform_conv_cuda.cpp:262-489).  The two dense contractions of the backward pass (W^T x gradOut
// and gradOut x columns^T) are plain library GEMMs issued by the host wrapper (s2anet_b200/dcn.py), as in the
// reference; the kernels here do the sampling side:
//   * s2a_deform_im2col_f32: x, offset -> columns [B, C*kH*kW, Ho*Wo] (bilinear samples, K index c*kH*kW + i*kW + j);
//   * s2a_deform_col2im_f32: grad columns -> grad_input (bilinear scatter, atomicAdd) and grad_offset (one pass:
//     the reference runs two kernels and a 5x5 neighbourhood search per element; the four corners are known).
// One thread owns one output position and walks channels, so column reads/writes are coalesced over positions.
#include "common.cuh"

namespace s2a {

struct DcGeom {
  int B, C, H, W, Ho, Wo, kH, kW, sH, sW, pH, pW, dH, dW, dgroups;
};

struct Corner {
  int o00;               // offset of the top-left corner inside one channel plane (may be out of range: see flags)
  float hy, hx, ly, lx;  // bilinear fractions
  bool in, t, b, l, r;   // sample inside (-1, H) x (-1, W); corner rows / columns inside the map
};

// deform_conv_cuda_kernel.cu:210-228 (position) and :83-114 (bilinear corners)
__device__ __forceinline__ Corner make_corner(const DcGeom& g, const float* off_b, int dg, int tap, int ho, int wo) {
  const int i = tap / g.kW, j = tap - i * g.kW;
  const long long plane = (long long)g.Ho * g.Wo;
  const float* op = off_b + ((long long)dg * 2 * g.kH * g.kW + 2 * tap) * plane + (long long)ho * g.Wo + wo;
  const float h = (float)(ho * g.sH - g.pH + i * g.dH) + op[0];
  const float w = (float)(wo * g.sW - g.pW + j * g.dW) + op[plane];
  Corner c;
  c.in = h > -1.0f && w > -1.0f && h < (float)g.H && w < (float)g.W;
  const float hf = floorf(h), wf = floorf(w);
  const int y0 = (int)hf, x0 = (int)wf;
  c.ly = h - hf; c.lx = w - wf; c.hy = 1.0f - c.ly; c.hx = 1.0f - c.lx;
  c.t = y0 >= 0; c.b = y0 + 1 <= g.H - 1; c.l = x0 >= 0; c.r = x0 + 1 <= g.W - 1;
  c.o00 = y0 * g.W + x0;
  return c;
}

constexpr int kDcThreads = 128;
constexpr int kDcMaxTaps = 49;      // up to 7x7 kernels

__global__ void __launch_bounds__(kDcThreads)
deform_im2col_kernel(const float* __restrict__ x, const float* __restrict__ offset, float* __restrict__ col, DcGeom g,
                     int c_chunk) {
  const int pos = blockIdx.x * kDcThreads + threadIdx.x;
  const int b = blockIdx.z, c0 = blockIdx.y * c_chunk;
  const int HoWo = g.Ho * g.Wo, kk = g.kH * g.kW;
  if (pos >= HoWo) return;
  const int ho = pos / g.Wo, wo = pos - ho * g.Wo;
  const float* off_b = offset + (long long)b * g.dgroups * 2 * kk * HoWo;
  const int cpg = g.C / g.dgroups;
  for (int tap = 0; tap < kk; ++tap) {
    int dg_prev = -1;
    Corner cn{};
    for (int c = c0; c < min(c0 + c_chunk, g.C); ++c) {
      const int dg = c / cpg;
      if (dg != dg_prev) { cn = make_corner(g, off_b, dg, tap, ho, wo); dg_prev = dg; }
      float v = 0.0f;
      if (cn.in) {
        const float* p = x + ((long long)b * g.C + c) * g.H * g.W + cn.o00;
        const float v1 = (cn.t && cn.l) ? p[0] : 0.0f, v2 = (cn.t && cn.r) ? p[1] : 0.0f;
        const float v3 = (cn.b && cn.l) ? p[g.W] : 0.0f, v4 = (cn.b && cn.r) ? p[g.W + 1] : 0.0f;
        v = cn.hy * cn.hx * v1 + cn.hy * cn.lx * v2 + cn.ly * cn.hx * v3 + cn.ly * cn.lx * v4;
      }
      col[(((long long)b * g.C + c) * kk + tap) * HoWo + pos] = v;
    }
  }
}

__global__ void __launch_bounds__(kDcThreads)
deform_col2im_kernel(const float* __restrict__ gcol, const float* __restrict__ x, const float* __restrict__ offset,
                     float* __restrict__ grad_x, float* __restrict__ grad_offset, DcGeom g) {
  const int pos = blockIdx.x * kDcThreads + threadIdx.x;
  const int b = blockIdx.z, dg = blockIdx.y;
  const int HoWo = g.Ho * g.Wo, kk = g.kH * g.kW;
  if (pos >= HoWo) return;
  const int ho = pos / g.Wo, wo = pos - ho * g.Wo;
  const float* off_b = offset + (long long)b * g.dgroups * 2 * kk * HoWo;
  const int cpg = g.C / g.dgroups;
  for (int tap = 0; tap < kk; ++tap) {
    const Corner cn = make_corner(g, off_b, dg, tap, ho, wo);
    float gy = 0.0f, gx = 0.0f;
    if (cn.in) {
      const float w1 = cn.hy * cn.hx, w2 = cn.hy * cn.lx, w3 = cn.ly * cn.hx, w4 = cn.ly * cn.lx;
      for (int c = dg * cpg; c < (dg + 1) * cpg; ++c) {
        const float gv = gcol[(((long long)b * g.C + c) * kk + tap) * HoWo + pos];
        const long long plane = ((long long)b * g.C + c) * g.H * g.W + cn.o00;
        if (grad_x) {
          // deformable_col2im (:278-351): the bilinear scatter of the column gradient
          float* q = grad_x + plane;
          if (cn.t && cn.l) atomicAdd(q, w1 * gv);
          if (cn.t && cn.r) atomicAdd(q + 1, w2 * gv);
          if (cn.b && cn.l) atomicAdd(q + g.W, w3 * gv);
          if (cn.b && cn.r) atomicAdd(q + g.W + 1, w4 * gv);
        }
        if (grad_offset) {
          // get_coordinate_weight (:147-187): d(sample)/d(offset_h) and d(sample)/d(offset_w)
          const float* p = x + plane;
          const float v1 = (cn.t && cn.l) ? p[0] : 0.0f, v2 = (cn.t && cn.r) ? p[1] : 0.0f;
          const float v3 = (cn.b && cn.l) ? p[g.W] : 0.0f, v4 = (cn.b && cn.r) ? p[g.W + 1] : 0.0f;
          gy += gv * (cn.hx * (v3 - v1) + cn.lx * (v4 - v2));
          gx += gv * (cn.hy * (v2 - v1) + cn.ly * (v4 - v3));
        }
      }
    }
    if (grad_offset) {
      float* go = grad_offset + (((long long)b * g.dgroups + dg) * 2 * kk + 2 * tap) * HoWo + pos;
      go[0] += gy;            // the reference accumulates into the caller's (pre-zeroed) tensors
      go[HoWo] += gx;
    }
  }
}

static int fill_geom(DcGeom& g, int B, int C, int H, int W, int kH, int kW, int sH, int sW, int pH, int pW, int dH, int dW,
                     int dgroups) {
  S2A_CHECK_ARG(B >= 0 && C > 0 && H > 0 && W > 0 && kH > 0 && kW > 0 && kH * kW <= kDcMaxTaps, "deform conv: bad sizes");
  S2A_CHECK_ARG(sH > 0 && sW > 0 && dH > 0 && dW > 0 && pH >= 0 && pW >= 0, "deform conv: bad stride / dilation / padding");
  S2A_CHECK_ARG(dgroups > 0 && C % dgroups == 0, "deform conv: channels must be divisible by deformable groups");
  g.B = B; g.C = C; g.H = H; g.W = W; g.kH = kH; g.kW = kW; g.sH = sH; g.sW = sW; g.pH = pH; g.pW = pW; g.dH = dH; g.dW = dW;
  g.dgroups = dgroups;
  g.Ho = (H + 2 * pH - (dH * (kH - 1) + 1)) / sH + 1;
  g.Wo = (W + 2 * pW - (dW * (kW - 1) + 1)) / sW + 1;
  S2A_CHECK_ARG(g.Ho > 0 && g.Wo > 0, "deform conv: empty output");
  S2A_CHECK_ARG((long long)C * H * W < (1ll << 31) && B <= 65535, "deform conv: tensor too large for this kernel");
  return S2A_OK;
}

}  // namespace s2a

extern "C" int s2a_deform_im2col_f32(const float* x, const float* offset, float* columns, int B, int C, int H, int W, int kH,
                                     int kW, int strideH, int strideW, int padH, int padW, int dilH, int dilW, int dgroups,
                                     void* stream) {
  using namespace s2a;
  DcGeom g;
  if (int rc = fill_geom(g, B, C, H, W, kH, kW, strideH, strideW, padH, padW, dilH, dilW, dgroups)) return rc;
  if (B == 0) return S2A_OK;
  S2A_CHECK_ARG(x && offset && columns, "deform_im2col: null pointer");
  const int c_chunk = 32;
  dim3 grid((unsigned)ceil_div((int64_t)g.Ho * g.Wo, kDcThreads), (unsigned)ceil_div(C, c_chunk), (unsigned)B);
  deform_im2col_kernel<<<grid, kDcThreads, 0, (cudaStream_t)stream>>>(x, offset, columns, g, c_chunk);
  S2A_LAUNCH_OK("deform_im2col_kernel");
  return S2A_OK;
}

extern "C" int s2a_deform_col2im_f32(const float* grad_columns, const float* x, const float* offset, float* grad_input,
                                     float* grad_offset, int B, int C, int H, int W, int kH, int kW, int strideH, int strideW,
                                     int padH, int padW, int dilH, int dilW, int dgroups, void* stream) {
  using namespace s2a;
  DcGeom g;
  if (int rc = fill_geom(g, B, C, H, W, kH, kW, strideH, strideW, padH, padW, dilH, dilW, dgroups)) return rc;
  if (B == 0) return S2A_OK;
  S2A_CHECK_ARG(grad_columns && x && offset && (grad_input || grad_offset), "deform_col2im: null pointer");
  dim3 grid((unsigned)ceil_div((int64_t)g.Ho * g.Wo, kDcThreads), (unsigned)dgroups, (unsigned)B);
  deform_col2im_kernel<<<grid, kDcThreads, 0, (cudaStream_t)stream>>>(grad_columns, x, offset, grad_input, grad_offset, g);
  S2A_LAUNCH_OK("deform_col2im_kernel");
  return S2A_OK;
}
