// common.cuh -- status codes, error plumbing and small helpers shared by the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/s2a_b200.h"

namespace s2a {

void set_error(const char* fmt, ...);

#define S2A_CHECK_ARG(cond, ...)                  \
  do {                                            \
    if (!(cond)) {                                \
      s2a::set_error(__VA_ARGS__);                \
      return S2A_ERR_INVALID_ARGUMENT;            \
    }                                             \
  } while (0)

#define S2A_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      s2a::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return S2A_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

#define S2A_LAUNCH_OK(name)                                                            \
  do {                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) {                                                          \
      s2a::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));        \
      return S2A_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int sm_count();   // cached multiprocessor count of the current device

}  // namespace s2a
