// conv_f32.cu -- fp32 fused gather + implicit-GEMM convolution for sm_100a (the exact-arithmetic
// path of AlignConv / DeformConv / ORConv2d; the bf16 tensor-core path lives in conv_tc.cu).
//
// Replaces (reference):
//   * deform_conv_forward_cuda: models/dcn/src/deform_conv_cuda.cpp:152-260, with the im2col of
//     models/dcn/src/deform_conv_cuda_kernel.cu:189-242 and the bilinear rule of :83-114;
//   * AlignConv.get_offset + forward: models/alignconv.py:29-98 (offset field generated in-kernel
//     from the anchors, never written to memory);
//   * ORConv2d.forward: models/orn/modules/ORConv.py:77-82 = conv2d(x, ARF(weight), bias), with the
//     ARF rotation (models/orn/src/cuda/ActiveRotatingFilter_cuda.cu:19-46) folded into the weight
//     tile load and RotationInvariantPooling (models/orn/functions/rotation_invariant_pooling.py
//     :19-27) folded into the epilogue.
//
// One CTA computes a 64 (positions) x 64 (output channels) tile.  The reference materialises a
// [C*9, H*W] column matrix in HBM (151 MB at P3), runs cuBLAS over it, then transposes/copies the
// result; here the sampled A tile only ever exists in shared memory: per 8-channel chunk the CTA
// blends the four bilinear corners straight into smem, multiplies against the matching weight slice
// and keeps the 4x4 per-thread accumulators in registers until the fused ReLU/bias/pool epilogue.
#include "common.cuh"

namespace s2a {

// Output tile BM positions x BN channels per CTA of 16 x 16 threads: 64 x 64 (4 x 4 outputs per thread) or, for
// ungrouped convs with >= 128 output channels, 128 x 128 (8 x 8 per thread): 16 FMAs per 16-byte shared-memory load
// instead of 8, and every gathered A value reused by twice as many output channels.
constexpr int CK = 8;
constexpr int kConvThreads = 256;
constexpr int kMaxTaps = 9;           // kH*kW <= 9 on this path (3x3 and smaller)
constexpr int kNarrowBN = 64;         // the narrow tile's channel extent (groups > 1 need (Co/groups) % 64 == 0)

enum { MODE_DEFORM = 0, MODE_ALIGN = 1, MODE_PLAIN = 2 };

struct ConvParams {
  const float* x; const float* aux; const float* w; const float* bias; const uint8_t* arf_idx;
  float* out; float* pooled;
  int B, C, H, W, Co, Ho, Wo, kH, kW, sH, sW, pH, pW, dH, dW, groups, dgroups, relu;
  float astride;                       // AlignConv feature stride
  int arfO, arfI, nOri, nRot, pool;    // ORConv
};

struct Sample { int off[4]; float wt[4]; };   // corner offsets inside one channel plane + weights

// bilinear corner table entry following deform_conv_cuda_kernel.cu:83-114 / :228
__device__ __forceinline__ void make_sample(float h, float w, int H, int W, Sample& s) {
#pragma unroll
  for (int q = 0; q < 4; ++q) { s.off[q] = 0; s.wt[q] = 0.0f; }
  if (!(h > -1.0f && w > -1.0f && h < (float)H && w < (float)W)) return;
  const int hl = (int)floorf(h), wl = (int)floorf(w);
  const int hh = hl + 1, wh = wl + 1;
  const float lh = h - (float)hl, lw = w - (float)wl;
  const float uh = 1.0f - lh, uw = 1.0f - lw;
  if (hl >= 0 && wl >= 0) { s.off[0] = hl * W + wl; s.wt[0] = uh * uw; }
  if (hl >= 0 && wh <= W - 1) { s.off[1] = hl * W + wh; s.wt[1] = uh * lw; }
  if (hh <= H - 1 && wl >= 0) { s.off[2] = hh * W + wl; s.wt[2] = lh * uw; }
  if (hh <= H - 1 && wh <= W - 1) { s.off[3] = hh * W + wh; s.wt[3] = lh * lw; }
}

template <int MODE, int BM, int BN>
__global__ void __launch_bounds__(kConvThreads)
conv_gather_f32_kernel(const ConvParams p) {
  constexpr int BNP = BN + 4;            // padded B-tile row (keeps float4 alignment, breaks store conflicts)
  constexpr int RM = BM / 16, RN = BN / 16;      // outputs per thread: RM positions x RN channels
  static_assert((BM == 64 || BM == 128) && (BN == 64 || BN == 128 || BN == 256) && RM * RN <= 64, "tile shapes");
  extern __shared__ __align__(16) unsigned char s_raw[];
  Sample* s_samp = reinterpret_cast<Sample*>(s_raw);                                    // [BM*9]   18 KB
  float (*s_a)[BM] = reinterpret_cast<float (*)[BM]>(s_raw + sizeof(Sample) * BM * kMaxTaps);   // [72][64] 18 KB
  float (*s_b)[BNP] = reinterpret_cast<float (*)[BNP]>(s_raw + sizeof(Sample) * BM * kMaxTaps +
                                                       sizeof(float) * CK * kMaxTaps * BM);     // [72][68] 19 KB
  __shared__ uint8_t s_inv[8 * 72];                        // ARF inverse map [nRot][nEntry]
  __shared__ uint8_t s_ct[CK * kMaxTaps];                  // K row -> (channel in chunk << 4) | tap
  __shared__ uint16_t s_ok[BN];                            // ORConv: output channel -> (filter << 3) | rotation

  const int tid = threadIdx.x;
  const int taps = p.kH * p.kW;
  const int HoWo = p.Ho * p.Wo;
  const int b = blockIdx.z;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int Cg = p.C / p.groups, Cog = p.Co / p.groups;
  const int g = n0 / Cog;                                   // host guarantees tiles do not straddle groups
  const int cpd = p.C / p.dgroups;
  const float* xb = p.x + (size_t)b * p.C * p.H * p.W;

  if (MODE == MODE_PLAIN && p.arf_idx) {
    const int nEntry = p.nOri * taps;
    for (int i = tid; i < nEntry * p.nRot; i += kConvThreads) {
      const int l = i / p.nRot, k = i % p.nRot;
      s_inv[k * nEntry + ((int)p.arf_idx[i] - 1)] = (uint8_t)l;
    }
  }

  // ---- sample table (per position x tap); rebuilt per deformable group when dgroups > 1 ----
  auto build_samples = [&](int dg) {
    for (int i = tid; i < BM * taps; i += kConvThreads) {
      const int m = i / taps, t = i % taps;
      const int pos = m0 + m;
      Sample s;
#pragma unroll
      for (int q = 0; q < 4; ++q) { s.off[q] = 0; s.wt[q] = 0.0f; }
      if (pos < HoWo) {
        const int ho = pos / p.Wo, wo = pos % p.Wo;
        const int ti = t / p.kW, tj = t % p.kW;
        if (MODE == MODE_PLAIN) {
          const int h = ho * p.sH - p.pH + ti * p.dH, w = wo * p.sW - p.pW + tj * p.dW;
          if (h >= 0 && h < p.H && w >= 0 && w < p.W) { s.off[0] = h * p.W + w; s.wt[0] = 1.0f; }
        } else if (MODE == MODE_DEFORM) {
          const float* ob = p.aux + ((size_t)b * p.dgroups + dg) * 2 * taps * HoWo;
          const float oy = ob[(size_t)(2 * t) * HoWo + pos];
          const float ox = ob[(size_t)(2 * t + 1) * HoWo + pos];
          const float h = (float)(ho * p.sH - p.pH + ti * p.dH) + oy;
          const float w = (float)(wo * p.sW - p.pW + tj * p.dW) + ox;
          make_sample(h, w, p.H, p.W, s);
        } else {   // MODE_ALIGN: models/alignconv.py:29-86, operation order preserved
          const float* a = p.aux + ((size_t)b * HoWo + pos) * 5;
          const float ax = a[0] / p.astride, ay = a[1] / p.astride;
          const float aw = a[2] / p.astride, ah = a[3] / p.astride;
          const float cs = cosf(a[4]), sn = sinf(a[4]);
          const float dw = aw / 3.0f, dh = ah / 3.0f;
          const float fi = (float)(ti - 1), fj = (float)(tj - 1);
          const float tx = __fmul_rn(dw, fj), ty = __fmul_rn(dh, fi);
          const float xr = __fsub_rn(__fmul_rn(cs, tx), __fmul_rn(sn, ty));
          const float yr = __fadd_rn(__fmul_rn(sn, tx), __fmul_rn(cs, ty));
          const float xa = __fadd_rn(xr, ax), ya = __fadd_rn(yr, ay);
          const float offx = __fsub_rn(xa, __fadd_rn((float)wo, fj));
          const float offy = __fsub_rn(ya, __fadd_rn((float)ho, fi));
          // deform-conv adds the offset back onto the regular tap position (:223-227)
          const float h = __fadd_rn((float)(ho - 1 + ti), offy);
          const float w = __fadd_rn((float)(wo - 1 + tj), offx);
          make_sample(h, w, p.H, p.W, s);
        }
      }
      s_samp[m * kMaxTaps + t] = s;
    }
  };

  // 16 x 16 threads; thread (tx, ty) owns positions 64*q + 4*tx + i and channels 64*q + 4*ty + j (i, j < 4; q < RM/4, RN/4)
  const int tx = tid & 15, ty = tid >> 4;
  float acc[RM][RN];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = 0.0f;

  int cur_dg = -1;
  const int kk = CK * taps;                          // rows of the smem K chunk actually used
  const int HW = p.H * p.W, wrow = Cg * taps;
  // index tables: (c, t) of K row kr; (o, k) = (filter, rotation) of output channel n0 + nn (ORConv)
  for (int kr = tid; kr < kk; kr += kConvThreads) s_ct[kr] = (uint8_t)(((kr / taps) << 4) | (kr % taps));
  if (MODE == MODE_PLAIN && p.arf_idx)
    for (int nn = tid; nn < BN; nn += kConvThreads) s_ok[nn] = (uint16_t)((((n0 + nn) / p.nRot) << 3) | ((n0 + nn) % p.nRot));
  const int step_n = kConvThreads / kk, step_k = kConvThreads - step_n * kk;   // thread stride in (nn, kr)
  const int kr0 = tid % kk, nn0 = tid / kk;
  for (int c0 = 0; c0 < Cg; c0 += CK) {
    const int cin0 = g * Cg + c0;                    // first input channel of this chunk
    const int dg = (MODE == MODE_DEFORM) ? cin0 / cpd : 0;
    if (dg != cur_dg) {
      __syncthreads();
      build_samples(dg);
      cur_dg = dg;
    }
    __syncthreads();                                 // previous chunk consumed; samples visible
    // A tile: s_a[c*taps + t][m] = sum_q wt_q * x[cin0 + c][off_q].  A thread keeps its position m and walks the
    // K rows; (c, t) of a row come from a table -- the per-element divisions were half of this kernel's instructions.
    {
      const int m = tid & (BM - 1);
      const Sample* sm = s_samp + m * kMaxTaps;
      const float* xc = xb + (size_t)cin0 * HW;
#pragma unroll 4
      for (int kr = tid / BM; kr < kk; kr += kConvThreads / BM) {
        const int ct = s_ct[kr], c = ct >> 4, t = ct & 15;
        float v = 0.0f;
        if (c0 + c < Cg) {
          const Sample& s = sm[t];
          const float* plane = xc + c * HW;
          if (MODE == MODE_PLAIN) {
            v = s.wt[0] != 0.0f ? plane[s.off[0]] : 0.0f;
          } else {
            v = s.wt[0] * plane[s.off[0]] + s.wt[1] * plane[s.off[1]] + s.wt[2] * plane[s.off[2]] +
                s.wt[3] * plane[s.off[3]];
          }
        }
        s_a[kr][m] = v;
      }
    }
    // B tile: s_b[c*taps + t][n] = W[n0 + n][c0 + c][t]; consecutive threads read consecutive K rows of one filter
    // (contiguous in memory: (c0 + c) * taps + t = c0 * taps + row), the (row, n) pair advances incrementally
    {
      const float* wc = p.w + (size_t)c0 * taps;
      int kr = kr0, nn = nn0;
      while (nn < BN) {
        const int ct = s_ct[kr], c = ct >> 4, t = ct & 15;
        const int co = n0 + nn;
        float v = 0.0f;
        if (co < p.Co && c0 + c < Cg) {
          if (MODE == MODE_PLAIN && p.arf_idx) {
            // rotated filter (o, k): entry dst of input plane i comes from base entry inv[k][dst]
            const int ok = s_ok[nn], o = ok >> 3, k = ok & 7;
            const int cin = c0 + c;
            const int ii = p.nOri == 1 ? cin : cin / p.nOri, lay = p.nOri == 1 ? 0 : cin % p.nOri;
            const int nEntry = p.nOri * taps;
            const int l = s_inv[k * nEntry + lay * taps + t];
            v = p.w[((size_t)o * p.arfI + ii) * nEntry + l];
          } else {
            v = wc[(size_t)co * wrow + kr];
          }
        }
        s_b[kr][nn] = v;
        kr += step_k; nn += step_n;
        if (kr >= kk) { kr -= kk; ++nn; }
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kk; ++k) {
      float a4[RM], b4[RN];
#pragma unroll
      for (int q = 0; q < RM / 4; ++q) {
        const float4 av = *reinterpret_cast<const float4*>(&s_a[k][q * 64 + tx * 4]);
        a4[4 * q] = av.x; a4[4 * q + 1] = av.y; a4[4 * q + 2] = av.z; a4[4 * q + 3] = av.w;
      }
#pragma unroll
      for (int q = 0; q < RN / 4; ++q) {
        const float4 bv = *reinterpret_cast<const float4*>(&s_b[k][q * 64 + ty * 4]);
        b4[4 * q] = bv.x; b4[4 * q + 1] = bv.y; b4[4 * q + 2] = bv.z; b4[4 * q + 3] = bv.w;
      }
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
  }

  // ---- epilogue: bias, ReLU, NCHW store, optional orientation max-pool ----
#pragma unroll
  for (int j = 0; j < RN; ++j) {
    const int co = n0 + (j >> 2) * 64 + ty * 4 + (j & 3);
    const float bv = (p.bias && co < p.Co) ? p.bias[co] : 0.0f;
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      float v = acc[i][j] + bv;
      if (p.relu) v = fmaxf(v, 0.0f);
      acc[i][j] = v;
    }
    if (co < p.Co) {
      float* o = p.out + ((size_t)b * p.Co + co) * HoWo;
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        const int pos = m0 + (i >> 2) * 64 + tx * 4 + (i & 3);
        if (pos < HoWo) o[pos] = acc[i][j];
      }
    }
  }
  if (p.pooled) {
    // groups of `pool` (= 8) consecutive output channels: this thread's 4 + the partner's 4 (lane ^ 16)
#pragma unroll
    for (int qn = 0; qn < RN / 4; ++qn) {
      float mx[RM];
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        float v = fmaxf(fmaxf(acc[i][4 * qn], acc[i][4 * qn + 1]), fmaxf(acc[i][4 * qn + 2], acc[i][4 * qn + 3]));
        const float o = __shfl_xor_sync(0xffffffffu, v, 16);
        mx[i] = fmaxf(v, o);
      }
      const int cbase = n0 + qn * 64 + ty * 4;
      if ((ty & 1) == 0 && cbase < p.Co) {
        float* o = p.pooled + ((size_t)b * (p.Co / 8) + cbase / 8) * HoWo;
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          const int pos = m0 + (i >> 2) * 64 + tx * 4 + (i & 3);
          if (pos < HoWo) o[pos] = mx[i];
        }
      }
    }
  }
}

template <int BM, int BN>
constexpr size_t conv_smem_bytes() {
  return sizeof(Sample) * BM * kMaxTaps + sizeof(float) * CK * kMaxTaps * (BM + BN + 4);
}

template <int MODE, int BM, int BN>
static cudaError_t launch_conv_tile(const ConvParams& p, cudaStream_t st) {
  auto kern = conv_gather_f32_kernel<MODE, BM, BN>;
  constexpr size_t smem = conv_smem_bytes<BM, BN>();
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)ceil_div((int64_t)p.Ho * p.Wo, BM), (unsigned)ceil_div(p.Co, BN), (unsigned)p.B);
  kern<<<grid, kConvThreads, smem, st>>>(p);
  return cudaSuccess;
}

template <int MODE>
static cudaError_t launch_conv_mode(const ConvParams& p, cudaStream_t st) {
  // Wide tiles when a tile cannot straddle channel groups and there is enough work to fill the GPU with them.
  // (64 x 256 -- every gathered A value reused by all output channels -- was tried too: no faster for AlignConv and
  // slower for ORConv, whose weight-tile build pays the ARF index arithmetic per element; the build phases, which do
  // not overlap the FMA loop, are what bounds this kernel.)
  if (p.groups == 1 && p.Co >= 128 &&
      ceil_div((int64_t)p.Ho * p.Wo, 128) * ceil_div(p.Co, 128) * p.B >= sm_count())
    return launch_conv_tile<MODE, 128, 128>(p, st);
  return launch_conv_tile<MODE, 64, kNarrowBN>(p, st);
}

static int launch_conv(int mode, const ConvParams& p, cudaStream_t st) {
  if (mode == MODE_DEFORM) S2A_CUDA_OK(launch_conv_mode<MODE_DEFORM>(p, st));
  else if (mode == MODE_ALIGN) S2A_CUDA_OK(launch_conv_mode<MODE_ALIGN>(p, st));
  else S2A_CUDA_OK(launch_conv_mode<MODE_PLAIN>(p, st));
  S2A_LAUNCH_OK("conv_gather_f32_kernel");
  return S2A_OK;
}

}  // namespace s2a

extern "C" int s2a_deform_conv_forward_f32(const float* x, const float* offset, const float* weight, float* out,
                                           int B, int C, int H, int W, int Co, int kH, int kW, int strideH,
                                           int strideW, int padH, int padW, int dilH, int dilW, int groups,
                                           int dgroups, int relu, void* stream) {
  using namespace s2a;
  S2A_CHECK_ARG(B >= 0 && C > 0 && H > 0 && W > 0 && Co > 0, "deform_conv: bad tensor sizes");
  S2A_CHECK_ARG(kH > 0 && kW > 0, "kernel size should be greater than zero, but got kH: %d kW: %d", kH, kW);
  S2A_CHECK_ARG(strideH > 0 && strideW > 0, "stride should be greater than zero, but got dH: %d dW: %d", strideH,
                strideW);
  S2A_CHECK_ARG(dilH > 0 && dilW > 0, "dilation should be greater than 0, but got dilationH: %d dilationW: %d",
                dilH, dilW);
  S2A_CHECK_ARG(groups > 0 && dgroups > 0 && C % groups == 0 && Co % groups == 0, "deform_conv: bad groups");
  S2A_CHECK_ARG(C % dgroups == 0, "input channels must divide deformable group size");
  S2A_CHECK_ARG(H >= kH && W >= kW, "input image is smaller than kernel");
  const int Ho = (H + 2 * padH - (dilH * (kH - 1) + 1)) / strideH + 1;
  const int Wo = (W + 2 * padW - (dilW * (kW - 1) + 1)) / strideW + 1;
  S2A_CHECK_ARG(Ho >= 1 && Wo >= 1, "deform_conv: output size is too small");
  if (kH * kW > kMaxTaps) { set_error("deform_conv: kernels larger than 3x3 are not supported"); return S2A_ERR_UNSUPPORTED; }
  if (groups > 1 && (Co / groups) % kNarrowBN != 0) {
    set_error("deform_conv: groups > 1 needs (Co/groups) %% %d == 0", kNarrowBN);
    return S2A_ERR_UNSUPPORTED;
  }
  if (dgroups > 1 && ((C / dgroups) % CK != 0 || (C / groups) % CK != 0)) {
    set_error("deform_conv: deformable_groups > 1 needs channel blocks that are multiples of %d", CK);
    return S2A_ERR_UNSUPPORTED;
  }
  if (B == 0) return S2A_OK;
  S2A_CHECK_ARG(x && offset && weight && out, "deform_conv: null pointer");
  S2A_CHECK_ARG(B <= 65535, "deform_conv: batch must be <= 65535");
  ConvParams p{};
  p.x = x; p.aux = offset; p.w = weight; p.out = out;
  p.B = B; p.C = C; p.H = H; p.W = W; p.Co = Co; p.Ho = Ho; p.Wo = Wo; p.kH = kH; p.kW = kW;
  p.sH = strideH; p.sW = strideW; p.pH = padH; p.pW = padW; p.dH = dilH; p.dW = dilW;
  p.groups = groups; p.dgroups = dgroups; p.relu = relu;
  return launch_conv(MODE_DEFORM, p, (cudaStream_t)stream);
}

extern "C" int s2a_alignconv_forward_f32(const float* x, const float* anchors, const float* weight, float* out,
                                         int B, int C, int H, int W, int Co, float stride, void* stream) {
  using namespace s2a;
  S2A_CHECK_ARG(B >= 0 && C > 0 && H > 0 && W > 0 && Co > 0, "alignconv: bad tensor sizes");
  S2A_CHECK_ARG(stride > 0.0f, "alignconv: stride must be positive");
  if (B == 0) return S2A_OK;
  S2A_CHECK_ARG(x && anchors && weight && out, "alignconv: null pointer");
  S2A_CHECK_ARG(B <= 65535, "alignconv: batch must be <= 65535");
  ConvParams p{};
  p.x = x; p.aux = anchors; p.w = weight; p.out = out;
  p.B = B; p.C = C; p.H = H; p.W = W; p.Co = Co; p.Ho = H; p.Wo = W; p.kH = 3; p.kW = 3;
  p.sH = p.sW = 1; p.pH = p.pW = 1; p.dH = p.dW = 1; p.groups = 1; p.dgroups = 1; p.relu = 1;
  p.astride = stride;
  return launch_conv(MODE_ALIGN, p, (cudaStream_t)stream);
}

extern "C" int s2a_orconv_forward_f32(const float* x, const float* weight, const uint8_t* indices,
                                      const float* bias, float* out, float* pooled, int B, int H, int W, int O,
                                      int I, int nOri, int nRot, void* stream) {
  using namespace s2a;
  S2A_CHECK_ARG(B >= 0 && H > 0 && W > 0 && O > 0 && I > 0, "orconv: bad tensor sizes");
  S2A_CHECK_ARG(nOri >= 1 && nOri <= 8 && nRot >= 1 && nRot <= 8, "orconv: nOrientation/nRotation must be in [1, 8]");
  S2A_CHECK_ARG(!pooled || ((O * nRot) % 8 == 0), "orconv: pooling needs O*nRot %% 8 == 0");
  if (B == 0) return S2A_OK;
  S2A_CHECK_ARG(x && weight && indices && out, "orconv: null pointer");
  S2A_CHECK_ARG(B <= 65535, "orconv: batch must be <= 65535");
  ConvParams p{};
  p.x = x; p.w = weight; p.bias = bias; p.arf_idx = indices; p.out = out; p.pooled = pooled;
  p.B = B; p.C = I * nOri; p.H = H; p.W = W; p.Co = O * nRot; p.Ho = H; p.Wo = W; p.kH = 3; p.kW = 3;
  p.sH = p.sW = 1; p.pH = p.pW = 1; p.dH = p.dW = 1; p.groups = 1; p.dgroups = 1; p.relu = 0;
  p.arfO = O; p.arfI = I; p.nOri = nOri; p.nRot = nRot; p.pool = 8;
  return launch_conv(MODE_PLAIN, p, (cudaStream_t)stream);
}
