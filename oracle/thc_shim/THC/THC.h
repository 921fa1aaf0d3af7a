// oracle/thc_shim/THC/THC.h -- TEST INFRASTRUCTURE.  torch >= 1.11 no longer ships the THC headers the reference's
// ORN extension includes (models/orn/src/cuda/ActiveRotatingFilter_cuda.cu:5-7).  The sources only use three names from
// them -- THCudaCheck, THCCeilDiv and (through THCAtomics.cuh) atomicAdd overloads -- which this include directory maps
// onto their c10 / ATen successors so the UNMODIFIED reference .cu files compile for sm_100a (oracle/build_oracle.py).
#pragma once
#include <ATen/ceil_div.h>
#include <c10/cuda/CUDAException.h>
#define THCudaCheck(expr) C10_CUDA_CHECK(expr)
template <typename T> inline T THCCeilDiv(T a, T b) { return at::ceil_div(a, b); }
