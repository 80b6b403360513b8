// Shared helpers for libxeofs_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/xeofs_b200.h"

namespace xb {

void set_error(const char* fmt, ...);

#define XB_CHECK_ARG(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::xb::set_error(__VA_ARGS__);        \
      return XEOFS_E_INVALID;              \
    }                                      \
  } while (0)

#define XB_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::xb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return XEOFS_E_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define XB_LAUNCH_CHECK()                                                                       \
  do {                                                                                          \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess) {                                                                    \
      ::xb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return XEOFS_E_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t lpad(int64_t l) { return round_up(l, 16); }
static inline int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }

int num_sms();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// (x - p) with NaN -> 0 (all-NaN features and all-NaN samples contribute nothing; sanitizer.py:124)
__device__ __forceinline__ float shifted(float x, float p) {
  float v = x - p;
  return (v == v) ? v : 0.0f;
}

// streaming 128-bit load that does not pollute L1
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

}  // namespace xb
