// box_iou_rotated.cu -- pairwise rotated IoU, [n,5] x [m,5] -> [n,m], batched, for sm_100a.
//
// Replaces box_iou_rotated_cuda (reference: utils/box_iou_rotated/src/box_iou_rotated_cuda.cu
// :13-62 kernel, :65-101 host).  The reference gives one thread one pair and runs the full polygon
// clipper (2x24-point local arrays, double-precision sin/cos per pair) for every pair.
//
// Round-2 design (round 1: 64 x 256 tiles, CTA-wide lists and barriers between the phases, the clipper's 24-point
// arrays in local memory -- ncu: 26x the algorithmic DRAM reads, 13 of 32 lanes active in the clip phase):
//   * a CTA owns up to 256 rows (boxes1) x 256 columns (boxes2); after ONE barrier (per-box work: double-precision
//     sin/cos once per box, not per pair) its eight warps never synchronise with each other again.  Warp w owns
//     columns 32w .. 32w+31 (thread = column, its box summary in registers) and walks the rows;
//   * all-pairs pass: a ~20-flop branch-free test on 24-byte box summaries (rbox_fast_zero) proves ~95 % of the pairs
//     exactly zero; those are stored at once -- a warp writes 128 contiguous bytes of a row.  The rest goes to a
//     warp-private queue (ballot + popc, no atomics);
//   * whenever a queue holds 32 entries the warp runs the next stage on them with ALL lanes busy: the full
//     classification (area early-out, separating axes, collinearity guards) -> a second queue -> the clipper.
//     Partial rounds only happen once per tile, at the end;
//   * the clipper keeps its points in REGISTERS: candidates are staged through a per-thread shared-memory column and
//     come back as 16 named registers for a predicated sorting network + convexity check + fan area
//     (rbox_hull8, bit-identical to the reference's hull for 3..8 generic points; anything else -- more than eight
//     candidates, coincident points, a scan step that would pop -- falls back to the general 24-point routine).
// HBM traffic is the algorithmic minimum (4 B per pair out, 20 B per box in per tile); no local-memory traffic on
// the hot path.
#include <algorithm>

#include "common.cuh"
#include "rbox_iou.cuh"

namespace s2a {

constexpr int kIouThreads = 256;
constexpr int kIouWarps = kIouThreads / 32;
constexpr int kIouCols = 256;          // columns (boxes2) per CTA tile: 32 per warp
constexpr int kIouRowsMax = 256;       // rows (boxes1) per CTA tile (runtime: tile_rows <= this, a multiple of 32)
constexpr int kIouQueue = 64;          // entries per warp queue: < 32 left over + <= 32 appended per step

struct IouArgs {
  const float* boxes1; const float* boxes2; float* out;
  int64_t n, m, ld_out, out_batch_stride, row_begin, row_end;
  int tile_rows, tile_first, tile_step, compact, flags;
};

constexpr size_t iou_smem_bytes(int tile_rows) {
  return sizeof(RBox) * kIouCols + (sizeof(RBox) + sizeof(RFast) + sizeof(RAng)) * (size_t)tile_rows +
         2 * sizeof(uint16_t) * kIouQueue * kIouWarps + sizeof(float) * 16 * kIouThreads;
}

// the general 24-point clipper (thread-local arrays), out of line: reached by degenerate pairs only
__device__ __noinline__ float iou_clip_general(const RBox& A, const RBox& B) { return rbox_iou_clip(A, B); }

__global__ void __launch_bounds__(kIouThreads, 3)
box_iou_rotated_kernel(const IouArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  RBox* s_col = reinterpret_cast<RBox*>(smem);
  RBox* s_row = s_col + kIouCols;
  RFast* s_rowf = reinterpret_cast<RFast*>(s_row + a.tile_rows);
  RAng* s_rowa = reinterpret_cast<RAng*>(s_rowf + a.tile_rows);
  uint16_t* s_q = reinterpret_cast<uint16_t*>(s_rowa + a.tile_rows);
  float* s_pts = reinterpret_cast<float*>(s_q + 2 * kIouQueue * kIouWarps);

  const int tid = threadIdx.x, wid = tid >> 5;
  const unsigned lane = tid & 31, lt = (1u << lane) - 1u;
  const int64_t b = blockIdx.z;
  const int64_t tile = (int64_t)a.tile_first + (int64_t)blockIdx.x * a.tile_step;      // global row-tile index
  const int64_t row0 = a.row_begin + tile * a.tile_rows;
  const int64_t col0 = (int64_t)blockIdx.y * kIouCols;
  const int nr = (int)min((int64_t)a.tile_rows, a.row_end - row0);
  const int nc = (int)min((int64_t)kIouCols, a.m - col0);
  // compact: the output holds only the row tiles this launch computes, packed in launch order
  const int64_t orow0 = a.compact ? (int64_t)blockIdx.x * a.tile_rows : row0;
  float* o = a.out + b * a.out_batch_stride + orow0 * a.ld_out + col0;

  // ---- per-box work, once per tile: coalesced-ish 20-byte reads, one double-precision sin/cos per box ----
  RFast cf; RAng ca;
  cf.x = cf.y = cf.r = cf.mn = 0.0f; ca.s2t = ca.c2t = 0.0f;
  const bool col_ok = tid < nc;
  if (col_ok) {
    const float* g = a.boxes2 + (b * a.m + col0 + tid) * 5;
    RBox bx;
    rbox_prep(g[0], g[1], g[2], g[3], g[4], bx);
    s_col[tid] = bx;
    rbox_fast_of(bx, cf, ca);
  }
  for (int i = tid; i < nr; i += kIouThreads) {
    const float* g = a.boxes1 + (b * a.n + row0 + i) * 5;
    RBox bx;
    rbox_prep(g[0], g[1], g[2], g[3], g[4], bx);
    s_row[i] = bx;
    rbox_fast_of(bx, s_rowf[i], s_rowa[i]);
  }
  __syncthreads();                        // the only CTA-wide barrier

  uint16_t* qm = s_q + wid * kIouQueue;                                  // pairs the fast test could not decide
  uint16_t* qc = s_q + (kIouWarps + wid) * kIouQueue;                    // pairs to clip
  float* scratch = s_pts + tid;                                          // this thread's candidate column, stride 256
  const int cbase = wid * 32;
  const bool no_reject = (a.flags & S2A_IOU_NO_REJECT) != 0;
  int nm = 0, nq = 0;

  // One copy of every stage (instruction cache): the hot loop walks rows until the first queue holds a full round;
  // a stage runs when its queue has 32 entries -- all lanes busy -- or, once the rows are exhausted, to drain it.
  int r = 0;
  while (true) {
    // ---- all-pairs pass: thread = column, loop over the rows (row summaries are shared-memory broadcasts) ----
    for (; r < nr && nm < 32; ++r) {
      bool maybe = false;
      if (col_ok) {
        if (!no_reject && rbox_fast_zero(s_rowf[r], s_rowa[r], cf, ca)) o[(int64_t)r * a.ld_out + tid] = 0.0f;
        else maybe = true;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, maybe);
      if (bal) {
        if (maybe) qm[nm + __popc(bal & lt)] = (uint16_t)((r << 5) | lane);
        nm += __popc(bal);
        __syncwarp();
      }
    }
    const bool done = r >= nr;
    // ---- full classification of up to 32 queued pairs; survivors go to the clip queue ----
    if (nm >= 32 || (done && nm > 0)) {
      const int cnt = min(nm, 32);
      nm -= cnt;
      bool clip = false;
      int p = 0;
      if ((int)lane < cnt) {
        p = qm[nm + lane];
        const int pr = p >> 5, pc = cbase + (p & 31);
        const RBox& rb = s_row[pr];
        const RBox& cc = s_col[pc];
        int cls;
        if (no_reject) {
          const float a1 = RB_MUL(rb.w, rb.h), a2 = RB_MUL(cc.w, cc.h);
          cls = (a1 <= RB_LO_1E14 || a2 <= RB_LO_1E14) ? RB_ZERO : RB_CLIP;
        } else {
          cls = rbox_classify(rb, cc);
        }
        if (cls == RB_ZERO) o[(int64_t)pr * a.ld_out + pc] = 0.0f;
        else clip = true;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, clip);
      if (clip) qc[nq + __popc(bal & lt)] = (uint16_t)p;
      nq += __popc(bal);
      __syncwarp();
    }
    // ---- clip up to 32 queued pairs, one per lane ----
    if (nq >= 32 || (done && nm == 0 && nq > 0)) {
      const int cnt = min(nq, 32);
      nq -= cnt;
      if ((int)lane < cnt) {
        const int p = qc[nq + lane];
        const int pr = p >> 5, pc = cbase + (p & 31);
        bool ok;
        float v = rbox_iou_clip_try(s_row[pr], s_col[pc], scratch, kIouThreads, ok);
        if (!ok) v = iou_clip_general(s_row[pr], s_col[pc]);
        o[(int64_t)pr * a.ld_out + pc] = v;
      }
      __syncwarp();
    }
    if (done && nm == 0 && nq == 0) break;
  }
}

static int launch_iou(const float* boxes1, int64_t n, const float* boxes2, int64_t m, int64_t batch, float* out,
                      int64_t ld_out, int64_t out_batch_stride, int64_t row_begin, int64_t row_end, int tile_rows,
                      int tile_first, int tile_step, int compact, int flags, cudaStream_t st) {
  S2A_CHECK_ARG(n >= 0 && m >= 0 && batch >= 0, "box_iou_rotated: negative size (n=%lld m=%lld batch=%lld)",
                (long long)n, (long long)m, (long long)batch);
  S2A_CHECK_ARG(row_begin >= 0 && row_begin <= row_end && row_end <= n,
                "box_iou_rotated: row range [%lld, %lld) outside [0, %lld)", (long long)row_begin,
                (long long)row_end, (long long)n);
  S2A_CHECK_ARG(tile_step >= 1 && tile_first >= 0 && tile_first < tile_step,
                "box_iou_rotated: need 0 <= tile_first < tile_step (got %d, %d)", tile_first, tile_step);
  if (tile_rows <= 0) {
    // default: 256-row tiles; smaller ones while the grid would not fill the GPU twice over
    tile_rows = kIouRowsMax;
    const int64_t cols = ceil_div(std::max<int64_t>(m, 1), kIouCols) * std::max<int64_t>(batch, 1);
    while (tile_rows > 32 && ceil_div(row_end - row_begin, tile_rows) * cols < 2 * (int64_t)sm_count() && tile_step == 1)
      tile_rows >>= 1;
  }
  S2A_CHECK_ARG(tile_rows % 32 == 0 && tile_rows <= kIouRowsMax, "box_iou_rotated: tile_rows must be a multiple of 32 <= %d",
                kIouRowsMax);
  if (row_end == row_begin || m == 0 || batch == 0) return S2A_OK;
  S2A_CHECK_ARG(boxes1 && boxes2 && out, "box_iou_rotated: null pointer");
  S2A_CHECK_ARG(ld_out >= m, "box_iou_rotated: ld_out (%lld) < m (%lld)", (long long)ld_out, (long long)m);
  const int64_t ntiles = ceil_div(row_end - row_begin, tile_rows);
  const int64_t mine = tile_first < ntiles ? ceil_div(ntiles - tile_first, tile_step) : 0;     // tiles of this launch
  if (mine == 0) return S2A_OK;
  S2A_CHECK_ARG(batch <= 65535 && ceil_div(m, kIouCols) <= 65535 && mine < (1ll << 31),
                "box_iou_rotated: batch and ceil(m/256) must be <= 65535");
  if (out_batch_stride <= 0) out_batch_stride = (compact ? mine * tile_rows : n) * ld_out;
  IouArgs a{boxes1, boxes2, out, n, m, ld_out, out_batch_stride, row_begin, row_end, tile_rows, tile_first, tile_step,
            compact ? 1 : 0, flags};
  const size_t smem = iou_smem_bytes(tile_rows);
  S2A_CUDA_OK(cudaFuncSetAttribute(box_iou_rotated_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)iou_smem_bytes(kIouRowsMax)));
  dim3 grid((unsigned)mine, (unsigned)ceil_div(m, kIouCols), (unsigned)batch);
  box_iou_rotated_kernel<<<grid, kIouThreads, smem, st>>>(a);
  S2A_LAUNCH_OK("box_iou_rotated_kernel");
  return S2A_OK;
}

}  // namespace s2a

extern "C" int s2a_box_iou_rotated(const float* boxes1, int64_t n, const float* boxes2, int64_t m,
                                   int64_t batch, float* out, int64_t ld_out, int64_t row_begin,
                                   int64_t row_end, int flags, void* stream) {
  return s2a::launch_iou(boxes1, n, boxes2, m, batch, out, ld_out, 0, row_begin, row_end, 0, 0, 1, 0, flags,
                         (cudaStream_t)stream);
}

extern "C" int s2a_box_iou_rotated_tiles(const float* boxes1, int64_t n, const float* boxes2, int64_t m, int64_t batch,
                                         float* out, int64_t ld_out, int64_t out_batch_stride, int tile_rows,
                                         int tile_first, int tile_step, int compact, int flags, void* stream) {
  return s2a::launch_iou(boxes1, n, boxes2, m, batch, out, ld_out, out_batch_stride, 0, n, tile_rows, tile_first, tile_step,
                         compact, flags, (cudaStream_t)stream);
}
