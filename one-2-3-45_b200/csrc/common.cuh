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

// "Once per device" guard for per-function attributes (cudaFuncSetAttribute applies to the CURRENT device only, and a
// process may drive several devices, e.g. run.py --gpu_idx): need() is true the first time it is called on a device.
struct PerDeviceOnce {
  unsigned long long seen = 0;
  bool need() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d > 63) return true;
    if ((seen >> d) & 1ull) return false;
    seen |= 1ull << d;
    return true;
  }
};

// Number of SMs of the current device (cached).  Grids of persistent kernels are sized
// as a multiple of this (148 on B200).
int sm_count();

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Programmatic dependent launch (PDL).  The ~660 kernels of a UNet pass are launched with programmatic stream
// serialization: kernel N+1 may be scheduled while kernel N drains, runs its prologue (barrier init, TMEM allocation,
// descriptor prefetch, index math) and then blocks in pdl_wait() until kernel N has completed and flushed its writes.
// Every kernel launched this way calls pdl_wait() before its first access to memory another kernel may have written (or
// may still read), so the chain stays transitively ordered; kernels launched normally are unaffected (wait is a no-op).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();  // api.cu: environment O2345_PDL (default on; "0" switches the launch attribute off)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr, cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Same, for kernels that run as thread-block clusters of (cluster_x, 1, cluster_z) CTAs ((1, 1, 1): no cluster attribute).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                                      int cluster_z, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  int n = 1;
  if (cluster_x > 1 || cluster_z > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x, attr[n].val.clusterDim.y = 1, attr[n].val.clusterDim.z = cluster_z;
    ++n;
  }
  cfg.attrs = attr, cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace o2345
