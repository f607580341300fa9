// Shared helpers for the libp2c.so kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/point2cyl.h"

#define P2C_FULL_MASK 0xffffffffu

#define P2C_RETURN_IF_CUDA_ERROR()                 \
  do {                                             \
    cudaError_t e__ = cudaGetLastError();          \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

#define P2C_CUDA_TRY(expr)                         \
  do {                                             \
    cudaError_t e__ = (expr);                      \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

static inline int p2c_ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float p2c_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(P2C_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ double p2c_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(P2C_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ float p2c_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(P2C_FULL_MASK, v, o));
  return v;
}

// |v|^2 exactly as torch.sum(v ** 2, -1) rounds it: (v0^2 + v1^2) + v2^2, no contraction.
__device__ __forceinline__ float p2c_norm2_rn(float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}
// Expanded-form squared distance of models/pointnet_util.py:37-39 with torch.matmul's K=3 FMA chain.
// (ax,ay,az,na) is the `src` row, (bx,by,bz,nb) the `dst` row of square_distance(src, dst).
__device__ __forceinline__ float p2c_sqdist_expanded(float ax, float ay, float az, float na,
                                                     float bx, float by, float bz, float nb) {
  float dot = __fmaf_rn(az, bz, __fmaf_rn(ay, by, __fmul_rn(ax, bx)));
  return __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), na), nb);
}
