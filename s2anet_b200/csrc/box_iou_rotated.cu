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
constexpr int kIouRowsMax = 256;       // rows (boxes1) per CTA tile (runtime: tile_rows <= this, a multiple of 32)
// CPT = columns per thread: a warp owns 32 * CPT consecutive columns (boxes2), a CTA 256 * CPT.  Two columns per thread
// halve the per-pair share of the row-summary loads, the queue bookkeeping and the loop itself and give the
// scheduler two independent dependency chains; one column per thread wastes fewer lanes when m is small.
template <int CPT> __host__ __device__ constexpr int iou_cols() { return kIouThreads * CPT; }
template <int CPT> __host__ __device__ constexpr int iou_queue() { return 32 + 16 * 32 * CPT; }   // < 32 left over + <= 16 rows x 32 * CPT appended per block

struct IouArgs {
  const float* boxes1; const float* boxes2; float* out;
  int64_t n, m, ld_out, out_batch_stride, row_begin, row_end;
  int tile_rows;      // rows per CTA
  int deal_rows;      // rows per dealt tile (divides tile_rows; == tile_rows unless compact): the CTA's rows are
                      // tile_rows / deal_rows consecutive tiles of THIS launch, i.e. every tile_step-th global tile
  int tile_first, tile_step, compact, flags;
  int64_t mine_rows;  // valid rows of this launch in packed order (compact) -- bounds the last CTA
};

template <int CPT> constexpr size_t iou_smem_bytes(int tile_rows) {
  return sizeof(RBox) * iou_cols<CPT>() + (sizeof(RBox) + sizeof(RFast) + sizeof(RAng)) * (size_t)tile_rows +
         sizeof(uint16_t) * (iou_queue<CPT>() + 64) * kIouWarps + sizeof(float) * 16 * kIouThreads;
}

// the general 24-point clipper (thread-local arrays), out of line: reached by degenerate pairs only
__device__ __noinline__ float iou_clip_general(const RBox& A, const RBox& B) { return rbox_iou_clip(A, B); }

template <int CPT, bool NO_REJECT>
__global__ void __launch_bounds__(kIouThreads, 3)
box_iou_rotated_kernel(const IouArgs a) {
  constexpr int COLS = iou_cols<CPT>(), QN = iou_queue<CPT>(), WCOLS = 32 * CPT;
  constexpr int CBITS = CPT == 2 ? 6 : 5;                  // queue entry = row << CBITS | column inside the warp's group
  extern __shared__ __align__(16) uint8_t smem[];
  RBox* s_col = reinterpret_cast<RBox*>(smem);
  RBox* s_row = s_col + COLS;
  RFast* s_rowf = reinterpret_cast<RFast*>(s_row + a.tile_rows);
  RAng* s_rowa = reinterpret_cast<RAng*>(s_rowf + a.tile_rows);
  uint16_t* s_q = reinterpret_cast<uint16_t*>(s_rowa + a.tile_rows);
  float* s_pts = reinterpret_cast<float*>(s_q + (QN + 64) * kIouWarps);

  const int tid = threadIdx.x, wid = tid >> 5;
  const unsigned lane = tid & 31, lt = (1u << lane) - 1u;
  // grid = (batch, column tiles, row CTAs); CTAs are dispatched x-fastest, so all images' copies of a row CTA run
  // together, and the row CTAs are walked from the LAST one down: in anchor x GT matrices the last rows are the
  // large P5-P7 anchors whose pairs mostly reach the clipper (CTAs several times as long as the P3 ones), and the
  // tail of the launch should be made of short CTAs
  const int64_t b = blockIdx.x;
  const int64_t bx = (int64_t)gridDim.z - 1 - blockIdx.z;
  const int64_t col0 = (int64_t)blockIdx.y * COLS;
  const int nc = (int)min((int64_t)COLS, a.m - col0);
  // Rows of this CTA.  Dealt tiles are a.deal_rows high; CTA x takes the launch's tiles x*spt .. x*spt + spt - 1
  // (global tile = tile_first + local tile * tile_step).  Only the matrix's last tile can be partial and it is the
  // last tile of the launch that owns it, so the valid rows of a CTA are a prefix.
  const int spt = a.tile_rows / a.deal_rows;
  const int64_t lrow0 = bx * a.tile_rows;                  // first row in the launch's packed order
  const int nr = (int)min((int64_t)a.tile_rows, a.mine_rows - lrow0);
  auto global_row = [&](int i) -> int64_t {
    const int sub = i / a.deal_rows;
    const int64_t gt = (int64_t)a.tile_first + (bx * spt + sub) * a.tile_step;
    return a.row_begin + gt * a.deal_rows + (i - sub * a.deal_rows);
  };
  // compact: the output holds only the rows this launch computes, packed in launch order
  const int64_t orow0 = a.compact ? lrow0 : global_row(0);
  float* o = a.out + b * a.out_batch_stride + orow0 * a.ld_out + col0;

  // ---- per-box work, once per tile: 20-byte reads, one double-precision sin/cos per box ----
  const int cbase = wid * WCOLS;                           // first column of this warp
  RFast cf[CPT]; RAng ca[CPT];
  bool col_ok[CPT];
#pragma unroll
  for (int g = 0; g < CPT; ++g) {
    const int c = cbase + 32 * g + (int)lane;              // this thread's g-th column
    col_ok[g] = c < nc;
    cf[g].x = cf[g].y = cf[g].r = cf[g].mn = 0.0f; ca[g].s2t = ca[g].c2t = 0.0f;
    if (col_ok[g]) {
      const float* gp = a.boxes2 + (b * a.m + col0 + c) * 5;
      RBox bx;
      rbox_prep(gp[0], gp[1], gp[2], gp[3], gp[4], bx);
      s_col[c] = bx;
      rbox_fast_of(bx, cf[g], ca[g]);
    }
  }
  for (int i = tid; i < nr; i += kIouThreads) {
    const float* gp = a.boxes1 + (b * a.n + global_row(i)) * 5;
    RBox bx;
    rbox_prep(gp[0], gp[1], gp[2], gp[3], gp[4], bx);
    s_row[i] = bx;
    rbox_fast_of(bx, s_rowf[i], s_rowa[i]);
  }
  __syncthreads();                        // the only CTA-wide barrier
  if (cbase >= nc) return;                // (a warp whose columns are all past m)

  uint16_t* qm = s_q + wid * QN;                                         // pairs the fast test could not decide
  uint16_t* qc = s_q + kIouWarps * QN + wid * 64;                        // pairs to clip (< 32 left over + <= 32 per round)
  float* scratch = s_pts + tid;                                          // this thread's candidate column, stride 256
  int nm = 0, nq = 0;

  // One copy of every stage (instruction cache): the hot loop walks rows until the first queue holds a full round;
  // a stage runs when its queue has 32 entries -- all lanes busy -- or, once the rows are exhausted, to drain it.
  int r = 0;
  float* orow = o + cbase + lane;                          // this thread's first column in the current row
  const RFast* prf = s_rowf;
  const RAng* pra = s_rowa;
  while (true) {
    // ---- all-pairs pass: thread = column(s), loop over the rows (row summaries are shared-memory broadcasts).
    // Every element gets its 0.0f here, unconditionally -- a warp writes whole 128-byte lines, no sector is ever
    // written partially first (partial sectors cost DRAM fill reads on eviction) -- and the few pairs that turn
    // out to intersect are overwritten by the clip stage below.
    // Rows go in blocks of 16: the undecided pairs of a block are collected as one bit each in a per-thread register
    // (bit 16 g + row) and appended to the queue ONCE per block -- warp prefix sum of the popcounts, then every lane
    // writes its own few entries.  (A ballot / popc / store sequence per row ran on 96 % of the rows and took a
    // third of this loop's instructions; per 8 rows the divergent write loop still cost 13 % of the kernel: its trip
    // count is the MAXIMUM count over the lanes, which grows much more slowly than the block.)
    for (; r < nr && nm < 32; r += 16, orow += 16 * a.ld_out, prf += 16, pra += 16) {
      unsigned mk = 0u;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        unsigned sub = 0u;                                 // bit 8 g + row of this half block
        const int nrow = min(8, nr - r - 8 * h);
        const RFast* hf = prf + 8 * h;
        const RAng* ha = pra + 8 * h;
        float* ho = orow + (int64_t)(8 * h) * a.ld_out;
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          if (rr < nrow) {                                 // (warp-uniform; false only in the matrix's last tile)
            const RFast rf = hf[rr];
            const RAng ra = ha[rr];
#pragma unroll
            for (int g = 0; g < CPT; ++g) {
              const bool zero = !NO_REJECT && rbox_fast_zero(rf, ra, cf[g], ca[g]);
              if (col_ok[g]) {
                ho[(int64_t)rr * a.ld_out + 32 * g] = 0.0f;
                if (!zero) sub |= 1u << (g * 8 + rr);
              }
            }
          }
        }
        mk |= ((sub & 0xffu) << (8 * h)) | ((sub >> 8) << (16 + 8 * h));
      }
      // append: inclusive warp scan of the per-lane counts
      const int cntl = __popc(mk);
      int pos = cntl;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, pos, d);
        if ((int)lane >= d) pos += up;
      }
      const int total = __shfl_sync(0xffffffffu, pos, 31);
      if (total) {
        pos += nm - cntl;                                  // first slot of this lane
        while (mk) {
          const int bit = __ffs(mk) - 1;
          mk &= mk - 1u;
          qm[pos++] = (uint16_t)(((r + (bit & 15)) << CBITS) | ((bit >> 4) << 5) | lane);
        }
        nm += total;
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
        const int pr = p >> CBITS, pc = cbase + (p & (WCOLS - 1));
        const RBox& rb = s_row[pr];
        const RBox& cc = s_col[pc];
        int cls;
        if (NO_REJECT) {
          const float a1 = RB_MUL(rb.w, rb.h), a2 = RB_MUL(cc.w, cc.h);
          cls = (a1 <= RB_LO_1E14 || a2 <= RB_LO_1E14) ? RB_ZERO : RB_CLIP;
        } else {
          cls = rbox_classify(rb, cc);
        }
        clip = cls != RB_ZERO;                 // (zero: the element already holds its 0.0f)
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
        const int pr = p >> CBITS, pc = cbase + (p & (WCOLS - 1));
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

template <int CPT, bool NO_REJECT>
static int launch_iou_kernel(const IouArgs& a, int64_t mine, int64_t batch, cudaStream_t st) {
  auto kern = box_iou_rotated_kernel<CPT, NO_REJECT>;
  S2A_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)iou_smem_bytes<CPT>(kIouRowsMax)));
  dim3 grid((unsigned)batch, (unsigned)ceil_div(a.m, iou_cols<CPT>()), (unsigned)mine);
  kern<<<grid, kIouThreads, iou_smem_bytes<CPT>(a.tile_rows), st>>>(a);
  S2A_LAUNCH_OK("box_iou_rotated_kernel");
  return S2A_OK;
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
    // default: 256-row tiles (306 G pairs/s on config 4 against 299 / 274 with 128 / 64 rows), halved while the launch
    // would have fewer than ~8 waves of CTAs (a 256 x 512-pair CTA of large anchors runs > 100 us: the tail matters)
    tile_rows = kIouRowsMax;
    const int64_t cols = ceil_div(std::max<int64_t>(m, 1), iou_cols<2>()) * std::max<int64_t>(batch, 1);
    while (tile_rows > 32 && ceil_div(row_end - row_begin, tile_rows) * cols < 8ll * 3 * sm_count() && tile_step == 1)
      tile_rows >>= 1;
  }
  S2A_CHECK_ARG(tile_rows >= 32 && kIouRowsMax % tile_rows == 0, "box_iou_rotated: tile_rows must be 32, 64, 128 or %d",
                kIouRowsMax);
  if (row_end == row_begin || m == 0 || batch == 0) return S2A_OK;
  S2A_CHECK_ARG(boxes1 && boxes2 && out, "box_iou_rotated: null pointer");
  S2A_CHECK_ARG(ld_out >= m, "box_iou_rotated: ld_out (%lld) < m (%lld)", (long long)ld_out, (long long)m);
  const int64_t nrows = row_end - row_begin;
  const int64_t ntiles = ceil_div(nrows, tile_rows);
  const int64_t mine = tile_first < ntiles ? ceil_div(ntiles - tile_first, tile_step) : 0;     // tiles of this launch
  if (mine == 0) return S2A_OK;
  // valid rows of this launch: all its tiles are full except possibly the matrix's last one
  const int64_t last_tile = tile_first + (mine - 1) * tile_step;
  const int64_t mine_rows = (mine - 1) * tile_rows + std::min<int64_t>(tile_rows, nrows - last_tile * tile_rows);
  // packed output: CTAs are 256 rows high and take several dealt tiles each; in-place output: one dealt tile per CTA
  // (its rows must be consecutive in the output)
  // (64-row CTAs when 128-row ones would leave the launch with fewer than ~6 waves: the tail matters more than the
  // ~8 % of per-tile overhead)
  int cta_rows = tile_rows;
  if (compact) {
    cta_rows = kIouRowsMax;
    const int64_t waves8 = 6ll * 3 * sm_count();         // (6 waves: 3 % better than 8 for a quarter of config 4, equal elsewhere)
    while (cta_rows > std::max(tile_rows, 64) && ceil_div(mine_rows, cta_rows) * ceil_div(m, iou_cols<2>()) * batch < waves8)
      cta_rows >>= 1;
  }
  const int64_t nctas = ceil_div(mine_rows, cta_rows);
  S2A_CHECK_ARG(batch < (1ll << 31) && ceil_div(m, iou_cols<1>()) <= 65535 && nctas <= 65535,
                "box_iou_rotated: ceil(m/256) and the number of row CTAs must be <= 65535");
  if (out_batch_stride <= 0) out_batch_stride = (compact ? mine * tile_rows : n) * ld_out;
  IouArgs a{boxes1, boxes2, out, n, m, ld_out, out_batch_stride, row_begin, row_end, cta_rows, tile_rows, tile_first,
            tile_step, compact ? 1 : 0, flags, mine_rows};
  // two columns per thread unless one column per thread wastes fewer lanes on the last column tile
  const int64_t waste2 = ceil_div(m, iou_cols<2>()) * iou_cols<2>() - m, waste1 = ceil_div(m, iou_cols<1>()) * iou_cols<1>() - m;
  const bool two = waste2 <= waste1;
  const bool nr_ = (flags & S2A_IOU_NO_REJECT) != 0;
  if (two) return nr_ ? launch_iou_kernel<2, true>(a, nctas, batch, st) : launch_iou_kernel<2, false>(a, nctas, batch, st);
  return nr_ ? launch_iou_kernel<1, true>(a, nctas, batch, st) : launch_iou_kernel<1, false>(a, nctas, batch, st);
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
