// umma_bench2.cu -- microbenchmark 2: the conv_tc pipeline skeleton without any data movement.
// warp 0 lane 0: MMA issuer (waits full[s], issues 4 MMAs, commits empty[s]); warp 1 lane 0: "loader"
// (waits empty[s], arrives full[s]); optional epilogue warps 2-5 that tcgen05.ld the idle accumulator.
// Prints cycles per k-block (ideal 512).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma1(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit1(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) { unsigned n = 0; while (!try_wait(bar, parity)) { if (++n > (1u << 24)) { printf("timeout bar %u thread %d\n", bar, (int)threadIdx.x); __trap(); } } }
__device__ __forceinline__ void arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}

#ifndef RING
#define RING 4
#endif
constexpr int S = RING;   // barrier ring depth (operand buffers alias modulo 4)

__global__ void __launch_bounds__(192, 1) bench(int iters, int mode, long long* out, int N) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = sm;
  uint8_t* sB = sm + 4 * 16384;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 4 * 32768);     // full[S], empty[S], accfull[2], accempty[2]
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 4 * S + 5);
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u;
    h ^= h >> 13; h *= 2246822519u; h ^= h >> 16;
    reinterpret_cast<uint32_t*>(sm)[i] = (h & 0x807F807Fu) | 0x3F003F00u;
  }
  const uint32_t full = smem_u32(bar), empty = full + 8 * S, accfull = empty + 8 * S, accempty = accfull + 16, dummy = accempty + 16, full2 = dummy + 8, empty2 = full2 + 8 * S;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full + 8 * s) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty + 8 * s) : "memory");
    }
    for (int s = 0; s < 2; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(accfull + 8 * s) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(accempty + 8 * s) : "memory");
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(dummy) : "memory");
    for (int s = 0; s < S; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full2 + 8 * s) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty2 + 8 * s) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = 36;                        // k-blocks per tile
  const int tiles = iters / KB;
  if (warp == 0) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
      const long long c0 = clock64();
      int s = 0; uint32_t ph = 0;
      for (int t = 0; t < tiles; ++t) {
        const int as = t & 1;
        if (mode & 2) { wait(accempty + 8 * as, (((uint32_t)t >> 1) & 1u) ^ 1u); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
        if (mode & 16) {
          // software-pipelined: the wait for the NEXT stage sits between MMA 1 and MMA 2 of the current one
          if (t == 0) { wait(full + 8 * s, ph); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
          for (int kb = 0; kb < KB; ++kb) {
            const uint64_t ad = desc_sw128(smem_u32(sA + (s & 3) * 16384)), bd = desc_sw128(smem_u32(sB + (s & 3) * 32768));
            mma1(tm + as * 256, ad, bd, idesc, kb != 0);
            mma1(tm + as * 256, ad + 2, bd + 2, idesc, 1);
            int ns = s + 1; uint32_t nph = ph;
            if (ns == S) { ns = 0; nph ^= 1u; }
            if (!(t == tiles - 1 && kb == KB - 1)) { wait(full + 8 * ns, nph); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
            mma1(tm + as * 256, ad + 4, bd + 4, idesc, 1);
            mma1(tm + as * 256, ad + 6, bd + 6, idesc, 1);
            commit1(empty + 8 * s);
            s = ns; ph = nph;
          }
        } else
        for (int kb = 0; kb < KB; ++kb) {
          if (mode & 1) { wait(full + 8 * s, ph); if (mode & 8) wait(full2 + 8 * s, ph); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
          // cost of the pieces, without a partner thread: a try_wait that always succeeds (fresh barrier, parity 1),
          // a test_wait, a fence
          if (mode & 32) wait(accempty + 8, 1u);
          if (mode & 128) {
            uint32_t ok;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(accempty + 8), "r"(1u) : "memory");
            if (!ok) __trap();
          }
          if (mode & 64) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t ad = desc_sw128(smem_u32(sA + (s & 3) * 16384)), bd = desc_sw128(smem_u32(sB + (s & 3) * 32768));
#pragma unroll
          for (int k = 0; k < 4; ++k) mma1(tm + as * 256, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
          if (mode & 256) commit1(dummy);
          if (mode & 1) { commit1(empty + 8 * s); if (mode & 4) commit1(dummy); if (mode & 8) commit1(empty2 + 8 * s); }
          if (++s == S) { s = 0; ph ^= 1u; }
        }
        if (mode & 2) commit1(accfull + 8 * as);
      }
      commit1(accfull + 8);      // final drain marker (only meaningful when mode & 2 == 0)
      if (!(mode & 2)) wait(accfull + 8, 0);
      const long long c1 = clock64();
      if (blockIdx.x == 0) out[0] = c1 - c0;
    }
  } else if (warp == 1) {
    if (lane == 0 && (mode & 1)) {
      int s = 0; uint32_t ph = 0;
      
      for (int i = 0; i < tiles * KB; ++i) {
        wait(empty + 8 * s, ph ^ 1u);
        arrive(full + 8 * s);
        if (mode & 8) { wait(empty2 + 8 * s, ph ^ 1u); arrive(full2 + 8 * s); }
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  } else if (mode & 2) {
    const int quad = warp & 3;
    float acc = 0.f;
    for (int t = 0; t < tiles; ++t) {
      const int as = t & 1;
      wait(accfull + 8 * as, ((uint32_t)t >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(tm + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 256 + c0))
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(v[i]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      arrive(accempty + 8 * as);
    }
    if (acc == 123.456f) out[1] = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
  }
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 256;
  long long* d;
  cudaMalloc(&d, 16);
  const size_t smem = 1024 + 4 * 16384 + 4 * 32768 + 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 36 * 28;
  printf("ring depth %d, N = %d\n", S, N);
  for (int grid : {148})
    for (int mode : {0, 32, 64, 96, 128, 192, 256, 288, 352}) {
      for (int rep = 0; rep < 2; ++rep) {
        bench<<<grid, 192, smem>>>(iters, mode, d, N);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long cyc;
      cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("grid=%3d mode=%d (handshake=%d epilogue=%d 2commits=%d): %7.1f cycles per k-block (ideal 512)\n", grid, mode, mode & 1,
             (mode >> 1) & 1, (mode >> 2) & 1, (double)cyc / iters);
      fflush(stdout);
    }
  return 0;
}
