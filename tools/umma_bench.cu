// umma_bench.cu -- microbenchmark: back-to-back tcgen05.mma (kind::f16, bf16) issue rate on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/umma_bench tools/umma_bench.cu
// Prints cycles per MMA for N in {64,128,256}, cta_group 1/2, same-accumulator vs alternating accumulators,
// with 1 cluster and with all SMs busy.  Operand contents are irrelevant (zero-filled smem).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
template <int CG> __device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
template <int CG> __device__ __forceinline__ void commit(uint32_t bar) {
  if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}

template <int CG>
__global__ void __launch_bounds__(128, 1) bench(int N, int alt, int iters, int kstep_bytes, long long* out, int mode, int mper) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = sm;                  // 4 stages x 16 KB
  uint8_t* sB = sm + 4 * 16384;      // 4 stages x 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 4 * 32768);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 13; h *= 2246822519u; h ^= h >> 16;
    // two bf16 values in [-2, 2): sign + exponent 126..127 + random mantissa
    const uint32_t v = (h & 0x807F807Fu) | 0x3F003F00u | ((h >> 9) & 0x00800080u);
    reinterpret_cast<uint32_t*>(sm)[i] = (mode & 1) ? v : 0u;
  }
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(smem_u32(bar + 1)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tptr)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tptr)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *tptr;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((mper * CG) >> 4) << 24);
    const long long c0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int st = i & 3;
      const uint64_t ad = desc_sw128(smem_u32(sA + st * 16384)), bd = desc_sw128(smem_u32(sB + st * 32768));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t d = tm + (alt ? (uint32_t)(((i * 4 + k) & 1) * 256) : 0u);
        mma<CG>(d, ad + (uint64_t)(k * kstep_bytes >> 4), bd + (uint64_t)(k * kstep_bytes >> 4), idesc);
      }
      if (mode & 2) { commit<CG>(smem_u32(bar + 1)); commit<CG>(smem_u32(bar + 1)); }
    }
    commit<CG>(smem_u32(bar));
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    const long long c1 = clock64();
    if (blockIdx.x == 0) out[0] = c1 - c0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  else __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
  }
}

template <int CG> void run(int grid, int N, int alt, int kstep, int mode = 0, int mper = 128) {
  long long* d;
  cudaMalloc(&d, 8);
  const size_t smem = 1024 + 4 * 16384 + 4 * 32768 + 64;
  cudaFuncSetAttribute(bench<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, bench<CG>, N, alt, iters, kstep, d, mode, mper);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) { printf("error %s %s\n", cudaGetErrorString(e), cudaGetErrorString(e2)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    if (rep == 1) {
      const double per = (double)cyc / (iters * 4.0);
      const double flop = 2.0 * mper * CG * N * 16;              // per MMA (whole group)
      printf("mode=%d Mper=%d cg=%d grid=%3d N=%3d alt=%d kstep=%3dB : %7.1f cyc/MMA  -> %6.0f flop/clk/SM, %.3f ms, %7.0f TFLOP/s chip\n", mode, mper, CG, grid, N, alt,
             kstep, per, flop / per / CG, ms, flop * iters * 4.0 * (grid / CG) / (ms * 1e-3) / 1e12);
    }
  }
  cudaFree(d);
}

int main(int argc, char** argv) {
  if (argc > 2) {          // commit cost: two commits per 4 MMAs (mode 2) vs none, small and large N
    for (int N : {32, 256})
      for (int mode : {0, 2}) { run<1>(148, N, 0, 32, mode); run<2>(148, N, 0, 32, mode); }
    return 0;
  }
  if (argc > 1) {          // shape sweep: MMA time vs M (rows per CTA) and N
    for (int mper : {128, 64}) {
      for (int N : {16, 32, 64, 128, 256}) {
        run<1>(148, N, 0, 32, 0, mper);
        run<2>(148, N, 0, 32, 0, mper);
      }
    }
    return 0;
  }
  for (int grid : {2, 148}) {
    for (int N : {64, 128, 256}) {
      run<1>(grid, N, 0, 32);
      run<2>(grid, N, 0, 32);
    }
    run<1>(grid, 256, 1, 32);
    run<2>(grid, 256, 1, 32);
    run<1>(grid, 256, 0, 0);      // same K slice every time (A/B smem rows re-read)
    run<2>(grid, 256, 0, 0);
    for (int mode : {1, 2, 3}) {
      run<1>(grid, 256, 0, 32, mode);
      run<2>(grid, 256, 0, 32, mode);
    }
  }
  return 0;
}
