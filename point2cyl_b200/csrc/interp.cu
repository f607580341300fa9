// 3-NN inverse-distance feature interpolation (replaces models/pointnet_util.py:301-308).
//
// The reference materialises the (B,N,S) distance matrix and fully sorts it to take 3 entries.
// Here one CTA handles QPB query points of one cloud: the S source points sit in shared memory as
// SoA (x, y, z, |p|^2); phase 1 is one thread per query keeping a sorted top-3 in registers
// (strict '<' insertion, so equal distances keep the lower index first, like a stable sort);
// phase 2 is one warp per query doing the coalesced weighted gather of three D-float rows
// (float4 per lane) straight into the caller's (possibly strided, e.g. concat) output rows.
// Distances use the reference's expanded form, which can be slightly negative; not clamped.
//
// Algorithmic traffic per cloud: 12*N + 12*S + 4*S*D (source rows, L2 resident) + 4*N*D (output).
#include "common.cuh"

namespace {

constexpr int QPB = 128;

__global__ void __launch_bounds__(QPB)
three_nn_interp_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                       const float* __restrict__ feats2, int64_t ldf, int N, int S, int D,
                       float* __restrict__ out, int64_t ldo, int64_t* __restrict__ idx_out,
                       float* __restrict__ w_out) {
  extern __shared__ float4 sm4[];  // (x, y, z, |p|^2) per source point: one broadcast LDS.128 per candidate
  __shared__ int s_idx[QPB][3];
  __shared__ float s_w[QPB][3];
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const float* p2 = xyz2 + (size_t)b * S * 3;
  for (int i = tid; i < S; i += QPB) {
    float x = __ldg(p2 + i * 3), y = __ldg(p2 + i * 3 + 1), z = __ldg(p2 + i * 3 + 2);
    sm4[i] = make_float4(x, y, z, p2c_norm2_rn(x, y, z));
  }
  __syncthreads();

  const int q = blockIdx.x * QPB + tid;
  if (q < N) {
    const float* p1 = xyz1 + ((size_t)b * N + q) * 3;
    const float ax = __ldg(p1), ay = __ldg(p1 + 1), az = __ldg(p1 + 2);
    const float na = p2c_norm2_rn(ax, ay, az);
    float d0 = __int_as_float(0x7f800000), d1 = d0, d2 = d0;  // +inf
    int i0 = 0, i1 = 0, i2 = 0;
#pragma unroll 4
    for (int j = 0; j < S; ++j) {
      const float4 sp = sm4[j];
      const float d = p2c_sqdist_expanded(ax, ay, az, na, sp.x, sp.y, sp.z, sp.w);
      if (d < d2) {
        if (d < d1) {
          d2 = d1; i2 = i1;
          if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
          else { d1 = d; i1 = j; }
        } else { d2 = d; i2 = j; }
      }
    }
    // w = (1/(d+1e-8)) / sum, same op order as :305-307
    const float r0 = __fdiv_rn(1.0f, __fadd_rn(d0, 1e-8f));
    const float r1 = __fdiv_rn(1.0f, __fadd_rn(d1, 1e-8f));
    const float r2 = __fdiv_rn(1.0f, __fadd_rn(d2, 1e-8f));
    const float nrm = __fadd_rn(__fadd_rn(r0, r1), r2);
    const float w0 = __fdiv_rn(r0, nrm), w1 = __fdiv_rn(r1, nrm), w2 = __fdiv_rn(r2, nrm);
    s_idx[tid][0] = i0; s_idx[tid][1] = i1; s_idx[tid][2] = i2;
    s_w[tid][0] = w0; s_w[tid][1] = w1; s_w[tid][2] = w2;
    if (idx_out) {
      int64_t* io = idx_out + ((size_t)b * N + q) * 3;
      io[0] = i0; io[1] = i1; io[2] = i2;
    }
    if (w_out) {
      float* wo = w_out + ((size_t)b * N + q) * 3;
      wo[0] = w0; wo[1] = w1; wo[2] = w2;
    }
  }
  __syncthreads();

  const int lane = tid & 31, warp = tid >> 5;
  const float* f2 = feats2 + (size_t)b * S * ldf;
  const bool vec = (D % 4 == 0) && (ldf % 4 == 0) && (ldo % 4 == 0) &&
                   (((reinterpret_cast<uintptr_t>(feats2) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
  for (int t = warp; t < QPB; t += QPB / 32) {
    const int qq = blockIdx.x * QPB + t;
    if (qq >= N) break;
    const float* a0 = f2 + (size_t)s_idx[t][0] * ldf;
    const float* a1 = f2 + (size_t)s_idx[t][1] * ldf;
    const float* a2 = f2 + (size_t)s_idx[t][2] * ldf;
    const float w0 = s_w[t][0], w1 = s_w[t][1], w2 = s_w[t][2];
    float* o = out + ((size_t)b * N + qq) * ldo;
    if (vec) {
      for (int c = lane * 4; c < D; c += 128) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(a0 + c));
        const float4 v = __ldg(reinterpret_cast<const float4*>(a1 + c));
        const float4 w = __ldg(reinterpret_cast<const float4*>(a2 + c));
        float4 r;
        r.x = __fadd_rn(__fadd_rn(__fmul_rn(u.x, w0), __fmul_rn(v.x, w1)), __fmul_rn(w.x, w2));
        r.y = __fadd_rn(__fadd_rn(__fmul_rn(u.y, w0), __fmul_rn(v.y, w1)), __fmul_rn(w.y, w2));
        r.z = __fadd_rn(__fadd_rn(__fmul_rn(u.z, w0), __fmul_rn(v.z, w1)), __fmul_rn(w.z, w2));
        r.w = __fadd_rn(__fadd_rn(__fmul_rn(u.w, w0), __fmul_rn(v.w, w1)), __fmul_rn(w.w, w2));
        *reinterpret_cast<float4*>(o + c) = r;
      }
    } else {
      for (int c = lane; c < D; c += 32)
        o[c] = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(a0 + c), w0), __fmul_rn(__ldg(a1 + c), w1)),
                         __fmul_rn(__ldg(a2 + c), w2));
    }
  }
}

// S == 1: every point receives the single source row (models/pointnet_util.py:298-299).
__global__ void __launch_bounds__(256)
broadcast_rows_kernel(const float* __restrict__ feats2, int64_t ldf, int N, int D,
                      float* __restrict__ out, int64_t ldo, int64_t total) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / D;
    const int c = (int)(e - row * D);
    const int64_t b = row / N;
    out[row * ldo + c] = __ldg(feats2 + b * ldf + c);
  }
}

}  // namespace

extern "C" int p2c_three_nn_interp(const float* xyz1, const float* xyz2, const float* feats2,
                                   int64_t ldf, int B, int N, int S, int D, float* out, int64_t ldo,
                                   int64_t* idx_out, float* w_out, void* stream) {
  if (!feats2 || !out || B <= 0 || N <= 0 || S <= 0 || D <= 0 || ldf < D || ldo < D) return P2C_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (S == 1) {
    const int64_t total = (int64_t)B * N * D;
    const int blocks = (int)min((int64_t)148 * 16, (total + 255) / 256);
    broadcast_rows_kernel<<<blocks, 256, 0, st>>>(feats2, ldf, N, D, out, ldo, total);
    P2C_RETURN_IF_CUDA_ERROR();
    return 0;
  }
  if (!xyz1 || !xyz2) return P2C_EINVAL;
  if (S < 3) return P2C_EUNSUPPORTED;
  const size_t smem = (size_t)S * 4 * sizeof(float);
  if (smem > 200 * 1024) return P2C_EUNSUPPORTED;
  if (smem > 48 * 1024)
    P2C_CUDA_TRY(cudaFuncSetAttribute(three_nn_interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(p2c_ceil_div(N, QPB), B);
  three_nn_interp_kernel<<<grid, QPB, smem, st>>>(xyz1, xyz2, feats2, ldf, N, S, D, out, ldo, idx_out, w_out);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
