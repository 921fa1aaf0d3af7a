// orn.cu -- active rotating filters (forward / backward) and rotation-invariant pooling, sm_100a.
//
// Replaces (reference): ARF_forward_cuda / ARF_backward_cuda, models/orn/src/cuda/
// ActiveRotatingFilter_cuda.cu:19-46 + :79-119 and :48-76 + :122-163 (exported as arf_forward /
// arf_backward by models/orn/src/vision.cpp:7-12), and RotationInvariantPooling.forward,
// models/orn/functions/rotation_invariant_pooling.py:19-27.
//
// The reference scatters: one thread per SOURCE weight writes nRot destinations (uncoalesced
// stores).  Here one thread owns one DESTINATION element and reads its source through the inverted
// index table held in shared memory, so the 8x larger output is written with coalesced stores.
// These kernels keep the standalone arf_forward / arf_backward entry points of the reference alive;
// the fused ORConv2d kernels (conv_f32.cu, conv_tc.cu) apply the same inverse map while loading
// weights and never materialise the rotated bank.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <algorithm>

#include "common.cuh"

namespace s2a {

constexpr int kMaxEntry = 8 * 9;   // nOri * kH * kW
constexpr int kMaxRot = 8;

template <typename T>
__global__ void arf_forward_kernel(const T* __restrict__ w, const uint8_t* __restrict__ idx, T* __restrict__ out,
                                   int O, int I, int nEntry, int nRot) {
  __shared__ uint8_t s_inv[kMaxRot * kMaxEntry];
  for (int i = threadIdx.x; i < nEntry * nRot; i += blockDim.x) {
    const int l = i / nRot, k = i % nRot;
    s_inv[k * nEntry + ((int)idx[i] - 1)] = (uint8_t)l;
  }
  __syncthreads();
  const int64_t total = (int64_t)O * nRot * I * nEntry;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int dst = (int)(e % nEntry);
    const int64_t r = e / nEntry;
    const int i = (int)(r % I);
    const int64_t ok = r / I;
    const int k = (int)(ok % nRot);
    const int o = (int)(ok / nRot);
    out[e] = w[((int64_t)o * I + i) * nEntry + s_inv[k * nEntry + dst]];
  }
}

template <typename T>
__global__ void arf_backward_kernel(const T* __restrict__ gout, const uint8_t* __restrict__ idx, T* __restrict__ gw,
                                    int O, int I, int nEntry, int nRot) {
  __shared__ uint8_t s_idx[kMaxRot * kMaxEntry];
  for (int i = threadIdx.x; i < nEntry * nRot; i += blockDim.x) s_idx[i] = idx[i];
  __syncthreads();
  const int64_t total = (int64_t)O * I * nEntry;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)(e % nEntry);
    const int64_t r = e / nEntry;
    const int i = (int)(r % I);
    const int o = (int)(r / I);
    float acc = 0.0f;
    for (int k = 0; k < nRot; ++k) {
      const int src = (int)s_idx[l * nRot + k] - 1;
      acc += (float)gout[(((int64_t)o * nRot + k) * I + i) * nEntry + src];
    }
    gw[e] = (T)acc;
  }
}

template <typename T>
__device__ __forceinline__ float to_f(T v) { return (float)v; }

// x [B, C, HW] -> out [B, C/nOri, HW]
template <typename T>
__global__ void ri_pool_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t BG, int64_t HW, int nOri) {
  const int64_t total = BG * HW;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bg = e / HW, p = e % HW;
    const T* src = x + bg * nOri * HW + p;
    T best = src[0];
    float bf = to_f(best);
    for (int o = 1; o < nOri; ++o) {
      const T v = src[(int64_t)o * HW];
      const float vf = to_f(v);
      if (vf > bf) { bf = vf; best = v; }
    }
    out[e] = best;
  }
}

template <typename T>
static int arf_fwd_t(const void* w, const uint8_t* idx, void* out, int O, int I, int nEntry, int nRot, cudaStream_t st) {
  const int64_t total = (int64_t)O * nRot * I * nEntry;
  const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16);
  arf_forward_kernel<T><<<blocks, 256, 0, st>>>((const T*)w, idx, (T*)out, O, I, nEntry, nRot);
  S2A_LAUNCH_OK("arf_forward_kernel");
  return S2A_OK;
}
template <typename T>
static int arf_bwd_t(const void* g, const uint8_t* idx, void* gw, int O, int I, int nEntry, int nRot, cudaStream_t st) {
  const int64_t total = (int64_t)O * I * nEntry;
  const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16);
  arf_backward_kernel<T><<<blocks, 256, 0, st>>>((const T*)g, idx, (T*)gw, O, I, nEntry, nRot);
  S2A_LAUNCH_OK("arf_backward_kernel");
  return S2A_OK;
}
template <typename T>
static int ri_pool_t(const void* x, void* out, int64_t BG, int64_t HW, int nOri, cudaStream_t st) {
  const int64_t total = BG * HW;
  const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 32);
  ri_pool_kernel<T><<<blocks, 256, 0, st>>>((const T*)x, (T*)out, BG, HW, nOri);
  S2A_LAUNCH_OK("ri_pool_kernel");
  return S2A_OK;
}

static int check_arf(int O, int I, int nOri, int kH, int kW, int nRot) {
  S2A_CHECK_ARG(O >= 0 && I >= 0, "arf: negative plane count");
  S2A_CHECK_ARG(nOri >= 1 && kH >= 1 && kW >= 1 && nOri * kH * kW <= kMaxEntry,
                "arf: nOrientation*kH*kW must be in [1, %d]", kMaxEntry);
  S2A_CHECK_ARG(nRot >= 1 && nRot <= kMaxRot, "arf: nRotation must be in [1, %d]", kMaxRot);
  return S2A_OK;
}

}  // namespace s2a

extern "C" int s2a_arf_forward(const void* weight, const uint8_t* indices, void* out, int O, int I, int nOri, int kH,
                               int kW, int nRot, int dtype, void* stream) {
  using namespace s2a;
  int rc = check_arf(O, I, nOri, kH, kW, nRot);
  if (rc != S2A_OK) return rc;
  if ((int64_t)O * I == 0) return S2A_OK;
  S2A_CHECK_ARG(weight && indices && out, "arf_forward: null pointer");
  const int nEntry = nOri * kH * kW;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case S2A_F32: return arf_fwd_t<float>(weight, indices, out, O, I, nEntry, nRot, st);
    case S2A_BF16: return arf_fwd_t<__nv_bfloat16>(weight, indices, out, O, I, nEntry, nRot, st);
    case S2A_F16: return arf_fwd_t<__half>(weight, indices, out, O, I, nEntry, nRot, st);
  }
  set_error("arf_forward: unknown dtype %d", dtype);
  return S2A_ERR_INVALID_ARGUMENT;
}

extern "C" int s2a_arf_backward(const void* grad_out, const uint8_t* indices, void* grad_weight, int O, int I,
                                int nOri, int kH, int kW, int nRot, int dtype, void* stream) {
  using namespace s2a;
  int rc = check_arf(O, I, nOri, kH, kW, nRot);
  if (rc != S2A_OK) return rc;
  if ((int64_t)O * I == 0) return S2A_OK;
  S2A_CHECK_ARG(grad_out && indices && grad_weight, "arf_backward: null pointer");
  const int nEntry = nOri * kH * kW;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case S2A_F32: return arf_bwd_t<float>(grad_out, indices, grad_weight, O, I, nEntry, nRot, st);
    case S2A_BF16: return arf_bwd_t<__nv_bfloat16>(grad_out, indices, grad_weight, O, I, nEntry, nRot, st);
    case S2A_F16: return arf_bwd_t<__half>(grad_out, indices, grad_weight, O, I, nEntry, nRot, st);
  }
  set_error("arf_backward: unknown dtype %d", dtype);
  return S2A_ERR_INVALID_ARGUMENT;
}

extern "C" int s2a_ri_pool_forward(const void* x, void* out, int64_t B, int64_t C, int64_t HW, int nOri, int dtype,
                                   void* stream) {
  using namespace s2a;
  S2A_CHECK_ARG(B >= 0 && C >= 0 && HW >= 0, "ri_pool: negative size");
  S2A_CHECK_ARG(nOri >= 1 && C % nOri == 0, "ri_pool: channels (%lld) not divisible by nOrientation (%d)",
                (long long)C, nOri);
  if (B * C * HW == 0) return S2A_OK;
  S2A_CHECK_ARG(x && out, "ri_pool: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t BG = B * (C / nOri);
  switch (dtype) {
    case S2A_F32: return ri_pool_t<float>(x, out, BG, HW, nOri, st);
    case S2A_BF16: return ri_pool_t<__nv_bfloat16>(x, out, BG, HW, nOri, st);
    case S2A_F16: return ri_pool_t<__half>(x, out, BG, HW, nOri, st);
  }
  set_error("ri_pool: unknown dtype %d", dtype);
  return S2A_ERR_INVALID_ARGUMENT;
}
