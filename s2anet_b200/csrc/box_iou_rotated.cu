// box_iou_rotated.cu -- pairwise rotated IoU, [n,5] x [m,5] -> [n,m], batched, for sm_100a.
//
// Replaces box_iou_rotated_cuda (reference: utils/box_iou_rotated/src/box_iou_rotated_cuda.cu
// :13-62 kernel, :65-101 host).  The reference gives one thread one pair and runs the full polygon
// clipper (2x24-point local arrays, double-precision sin/cos per pair) for every pair.  Here a CTA
// owns a 64x256 tile of the IoU matrix and works in three phases:
//   1. prepare: the 64 row boxes and 256 column boxes are loaded with coalesced reads and their
//      sin/cos evaluated ONCE per box (not once per pair) into shared memory;
//   2. fast classify: every pair runs a branch-free ~25-flop test (rbox_classify_fast); the ~93 %
//      proven to have IoU == 0 are stored at once (a warp writes 128 contiguous bytes of a row), the
//      rest are appended to a per-CTA work list with a warp ballot + one shared atomic per warp;
//   3. full classify of the listed pairs (separating axes, collinearity guard), compacting in place;
//   4. clip: the ~3 % survivors are clipped densely and each result stored to its element.
// Splitting 2/3/4 matters because the longer tests would otherwise run for almost every warp with
// one or two active lanes (ncu, round 1: 69 % of issued instructions were that divergent tail).
// HBM traffic is the algorithmic minimum (4 B per pair out, 20 B per box in per tile).
#include "common.cuh"
#include "rbox_iou.cuh"

namespace s2a {

constexpr int kTileR = 64;             // rows (boxes1) per CTA tile
constexpr int kTileC = 256;            // columns (boxes2) per CTA tile
constexpr int kIouThreads = 256;

__global__ void __launch_bounds__(kIouThreads)
box_iou_rotated_kernel(const float* __restrict__ boxes1, int64_t n, const float* __restrict__ boxes2,
                       int64_t m, float* __restrict__ out, int64_t ld_out, int64_t row_begin,
                       int64_t row_end, int flags) {
  __shared__ RBox s_row[kTileR];
  __shared__ RBox s_col[kTileC];
  __shared__ __align__(16) uint16_t s_list[kTileR * kTileC];   // 64 x 256 pairs: 14-bit pair ids
  float* s_raw = reinterpret_cast<float*>(s_list);             // raw boxes live here only until they are prepared
  __shared__ int s_count;

  const int tid = threadIdx.x;
  const int64_t b = blockIdx.z;
  const int64_t row0 = row_begin + (int64_t)blockIdx.x * kTileR;
  const int64_t col0 = (int64_t)blockIdx.y * kTileC;
  const int nr = (int)min((int64_t)kTileR, row_end - row0);
  const int nc = (int)min((int64_t)kTileC, m - col0);
  const float* g1 = boxes1 + (b * n + row0) * 5;
  const float* g2 = boxes2 + (b * m + col0) * 5;

  // phase 1: coalesced raw loads, then one thread per box does the double-precision sin/cos
  for (int i = tid; i < nr * 5; i += kIouThreads) s_raw[i] = g1[i];
  for (int i = tid; i < nc * 5; i += kIouThreads) s_raw[kTileR * 5 + i] = g2[i];
  if (tid == 0) s_count = 0;
  __syncthreads();
  for (int i = tid; i < nr + nc; i += kIouThreads) {
    if (i < nr) {
      const float* r = s_raw + i * 5;
      rbox_prep(r[0], r[1], r[2], r[3], r[4], s_row[i]);
    } else {
      const float* r = s_raw + kTileR * 5 + (i - nr) * 5;
      rbox_prep(r[0], r[1], r[2], r[3], r[4], s_col[i - nr]);
    }
  }
  __syncthreads();

  // phase 2: fast classification of every pair.  Thread t owns column t of the tile (its box stays
  // in registers) and walks the 64 rows (row box = shared-memory broadcast).  ~93 % of the pairs are
  // proven zero by the branch-free fast test and stored right away -- a warp writes 32 consecutive
  // floats of one output row; the others are appended to the work list (warp ballot + one shared
  // atomic per warp), so the longer tests below never run with mostly idle lanes.
  const unsigned lane = tid & 31;
  const bool no_reject = (flags & S2A_IOU_NO_REJECT) != 0;
  float* o = out + (b * n + row0) * ld_out + col0;
  RBox cb;
  const bool col_ok = tid < nc;
  if (col_ok) cb = s_col[tid];
#pragma unroll 4
  for (int r = 0; r < nr; ++r) {
    bool maybe = false;
    if (col_ok) {
      if (!no_reject && rbox_classify_fast(s_row[r], cb) == RB_ZERO) o[(int64_t)r * ld_out + tid] = 0.0f;
      else maybe = true;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, maybe);
    if (bal) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_count, __popc(bal));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (maybe) s_list[base + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)(r * kTileC + tid);
    }
  }
  __syncthreads();

  // phase 3: full classification (area early-out, separating axes, collinearity guard) of the listed
  // pairs, compacting the survivors in place: round q reads entries [256q, 256q+256) and, after a
  // barrier, appends at positions below the number of entries processed so far.
  const int cnt_maybe = s_count;
  __syncthreads();
  if (tid == 0) s_count = 0;
  __syncthreads();
  for (int k0 = 0; k0 < cnt_maybe; k0 += kIouThreads) {
    const int k = k0 + tid;
    bool clip = false;
    int p = 0;
    if (k < cnt_maybe) {
      p = s_list[k];
      const RBox& rb = s_row[p >> 8];
      const RBox& cc = s_col[p & (kTileC - 1)];
      int cls;
      if (no_reject) {
        float a1 = RB_MUL(rb.w, rb.h), a2 = RB_MUL(cc.w, cc.h);
        cls = (a1 <= RB_LO_1E14 || a2 <= RB_LO_1E14) ? RB_ZERO : RB_CLIP;
      } else {
        cls = rbox_classify(rb, cc);
      }
      if (cls == RB_ZERO) o[(int64_t)(p >> 8) * ld_out + (p & (kTileC - 1))] = 0.0f;
      else clip = true;
    }
    __syncthreads();                       // every entry of this round has been read
    const unsigned bal = __ballot_sync(0xffffffffu, clip);
    if (bal) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_count, __popc(bal));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (clip) s_list[base + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)p;
    }
  }
  __syncthreads();

  // phase 4: dense clipping of the survivors (typically < 3 % of the tile)
  const int cnt = s_count;
  for (int k = tid; k < cnt; k += kIouThreads) {
    const int p = s_list[k];
    const int r = p >> 8, c = p & (kTileC - 1);
    o[(int64_t)r * ld_out + c] = rbox_iou_clip(s_row[r], s_col[c]);
  }
}

}  // namespace s2a

extern "C" int s2a_box_iou_rotated(const float* boxes1, int64_t n, const float* boxes2, int64_t m,
                                   int64_t batch, float* out, int64_t ld_out, int64_t row_begin,
                                   int64_t row_end, int flags, void* stream) {
  using namespace s2a;
  S2A_CHECK_ARG(n >= 0 && m >= 0 && batch >= 0, "box_iou_rotated: negative size (n=%lld m=%lld batch=%lld)",
                (long long)n, (long long)m, (long long)batch);
  S2A_CHECK_ARG(row_begin >= 0 && row_begin <= row_end && row_end <= n,
                "box_iou_rotated: row range [%lld, %lld) outside [0, %lld)", (long long)row_begin,
                (long long)row_end, (long long)n);
  if (row_end == row_begin || m == 0 || batch == 0) return S2A_OK;
  S2A_CHECK_ARG(boxes1 && boxes2 && out, "box_iou_rotated: null pointer");
  S2A_CHECK_ARG(ld_out >= m, "box_iou_rotated: ld_out (%lld) < m (%lld)", (long long)ld_out, (long long)m);
  S2A_CHECK_ARG(batch <= 65535 && ceil_div(m, kTileC) <= 65535,
                "box_iou_rotated: batch and ceil(m/256) must be <= 65535");
  dim3 grid((unsigned)ceil_div(row_end - row_begin, kTileR), (unsigned)ceil_div(m, kTileC), (unsigned)batch);
  box_iou_rotated_kernel<<<grid, kIouThreads, 0, (cudaStream_t)stream>>>(boxes1, n, boxes2, m, out, ld_out,
                                                                         row_begin, row_end, flags);
  S2A_LAUNCH_OK("box_iou_rotated_kernel");
  return S2A_OK;
}
