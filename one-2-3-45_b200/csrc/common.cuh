// Shared helpers for libo2345_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/o2345.h"

namespace o2345 {

void set_error(const char* fmt, ...);

#define O2345_CHECK_ARG(cond, msg)                          \
  do {                                                      \
    if (!(cond)) {                                          \
      ::o2345::set_error("%s: %s", __func__, msg);          \
      return O2345_EINVAL;                                  \
    }                                                       \
  } while (0)

#define O2345_CUDA(call)                                                            \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ::o2345::set_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(e__)); \
      return O2345_ECUDA;                                                           \
    }                                                                               \
  } while (0)

#define O2345_LAUNCH_CHECK() O2345_CUDA(cudaGetLastError())

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Number of SMs of the current device (cached).  Grids of persistent kernels are sized
// as a multiple of this (148 on B200).
int sm_count();

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace o2345
