// Shared helpers for libartic_sm100.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "artic.h"

namespace artic {

void set_error(const char* fmt, ...);

#define ARTIC_CHECK_ARG(cond, msg)                                     \
  do {                                                                 \
    if (!(cond)) {                                                     \
      artic::set_error("%s: %s (%s)", __func__, msg, #cond);           \
      return ARTIC_EINVAL;                                             \
    }                                                                  \
  } while (0)

#define ARTIC_LAUNCH_CHECK()                                           \
  do {                                                                 \
    cudaError_t e__ = cudaGetLastError();                              \
    if (e__ != cudaSuccess) {                                          \
      artic::set_error("%s: launch failed: %s", __func__, cudaGetErrorString(e__)); \
      return ARTIC_ECUDA;                                              \
    }                                                                  \
  } while (0)

template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void st_f(T* p, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

__device__ __forceinline__ int64_t seq_base(const artic_seq_t& s, int n) {
  return (int64_t)(n / s.n_inner) * s.s_outer + (int64_t)(n % s.n_inner) * s.s_inner;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum (blockDim.x multiple of 32, <= 1024). Result valid in thread 0.
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  if (w == 0) {
    v = (lane < (int)(blockDim.x >> 5)) ? smem32[lane] : 0.f;
    v = warp_sum(v);
  }
  return v;
}

// Which kernel family took a contraction (artic_path_counts): conv 0..3, wgrad 4..7
enum { PATH_CONV_TC = 0, PATH_CONV_TC_X3 = 1, PATH_CONV_GENERIC = 2, PATH_CONV_C1 = 3,
       PATH_WGRAD_TC = 4, PATH_WGRAD_TC_X3 = 5, PATH_WGRAD_GENERIC = 6, PATH_WGRAD_C1 = 7 };
extern long long g_path_counts[12];

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace artic
