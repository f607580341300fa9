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

// SMs the persistent (one CTA per SM) tensor-core kernels may occupy: p2c_set_sm_budget (version.cu).  A pipelined
// caller that runs the geometry stage of the next batch on a second stream leaves it some SMs this way.
extern int g_p2c_sm_budget;
static inline int p2c_sm_budget(int device_sms) {
  return (g_p2c_sm_budget > 0 && g_p2c_sm_budget < device_sms) ? g_p2c_sm_budget : device_sms;
}

// Programmatic dependent launch (sm_90+): a kernel launched through p2c_launch with g_p2c_pdl set may be scheduled while
// its predecessor in the stream is still draining - its CTAs take SMs as they free up, run their local prologue
// (barrier init, tensor-memory allocation, descriptor prefetch) and block in p2c_grid_dep_wait() until the predecessor
// has COMPLETED and its writes are visible; nothing produced or consumed by a predecessor may be touched before that call.
// p2c_grid_dep_launch() (issued after the wait, so that "the dependent runs" implies "my own predecessor is complete")
// lets the next kernel in the stream start the same way.  Both are no-ops for a kernel launched without the attribute.
extern int g_p2c_pdl;
__device__ __forceinline__ void p2c_grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void p2c_grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t p2c_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_p2c_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// "xyz-first" operand of the tcgen05 layer kernel (linear_tc.cu, p2c_sa_xyz_linear): the layer's input rows are not
// read from memory - row r = (b, s, j) is relu(bn0(W0 (xyz[b, idx[r]] - new_xyz[b, s]) + b0)), the first (xyz-only)
// conv of a set-abstraction level recomputed in the operand transform (three FMAs per element) instead of written by
// p2c_sa_first_layer and read back (2 x 268 MB per forward at B = 32 x N = 8192).
struct P2cXyzFirst {
  const float* xyz; const float* new_xyz; const int64_t* idx;
  const float* W0; int64_t ldw0; const float* b0;
  int N, S, ns;
  const double* moments;   // nine moments of the centred coordinates (p2c_group_moments) or NULL: when set, a pending
                           // train-mode BatchNorm of the first conv takes its sums from them in closed form
};

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

// Philox4x32-10 (Salmon et al., SC'11): 128 random bits from a 128-bit counter and a 64-bit key.  Used for the
// in-kernel dropout mask of the output heads: keep-bit of channel c of point m = bit (c & 31) of word (c >> 5) & 3 of
// philox(counter = {m, c >> 7, seed[1]}, key = seed[0]) - forward and backward regenerate the same bits from the
// two 64-bit seed words instead of reading a (B, C, N) float mask from HBM.
struct P2CPhilox4 { uint32_t v[4]; };
__device__ __forceinline__ P2CPhilox4 p2c_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  P2CPhilox4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}
// keep-bits of channels [128*blk, 128*blk + 128) of point m (p = 0.5: one bit per element, kept values scale by 2)
__device__ __forceinline__ P2CPhilox4 p2c_dropout_bits(const int64_t* seed, int64_t m, int blk) {
  const uint64_t s0 = (uint64_t)seed[0], s1 = (uint64_t)seed[1];
  return p2c_philox4x32_10((uint32_t)m, (uint32_t)((uint64_t)m >> 32) ^ ((uint32_t)blk << 24), (uint32_t)s1,
                           (uint32_t)(s1 >> 32), (uint32_t)s0, (uint32_t)(s0 >> 32));
}
__device__ __forceinline__ float p2c_dropout_scale(const P2CPhilox4& b, int c) {
  return ((b.v[(c >> 5) & 3] >> (c & 31)) & 1u) ? 2.0f : 0.0f;
}
