// Shared helpers for the EMAGE hot-path kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PM_OK 0
#define PM_EBADARG (-1)
#define PM_EUNSUPPORTED (-2)

// Every entry point: validate -> launch -> return cudaGetLastError() (positive) or PM_E* (negative).
#define PM_LAUNCH_CHECK()                                  \
  do {                                                     \
    cudaError_t _e = cudaGetLastError();                   \
    return _e == cudaSuccess ? PM_OK : (int)_e;            \
  } while (0)

#define PM_REQUIRE(cond) \
  do {                   \
    if (!(cond)) return PM_EBADARG; \
  } while (0)

static inline int pm_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

enum PmAct { PM_ACT_NONE = 0, PM_ACT_RELU = 1, PM_ACT_LEAKY = 2 };

__device__ __forceinline__ float pm_act(float v, int act, float slope) {
  if (act == PM_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == PM_ACT_LEAKY) return v > 0.f ? v : v * slope;
  return v;
}

__device__ __forceinline__ float pm_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float pm_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
