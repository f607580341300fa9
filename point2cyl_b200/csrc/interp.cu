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
                       float* __restrict__ w_out, const int64_t* __restrict__ idx_in, const float* __restrict__ w_in) {
  // three modes: search + gather (the fused call), search only (feats2 == NULL), gather only (idx_in != NULL)
  // sources in PAIRS: sA[p] = (x0, x1, y0, y1), sB[p] = (z0, z1, |p0|^2, |p1|^2) for sources 2p, 2p+1 - two broadcast
  // LDS.128 feed six packed FMUL2 / FFMA2 / FADD2 (same rounding per half as the scalar ops: bit-identical distances)
  extern __shared__ float4 sm4[];
  float4* sA = sm4;
  float4* sB = sm4 + (S + 1) / 2;
  __shared__ int s_idx[QPB][3];
  __shared__ float s_w[QPB][3];
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  if (idx_in) {
    const int q = blockIdx.x * QPB + tid;
    if (q < N) {
      const size_t o = ((size_t)b * N + q) * 3;
#pragma unroll
      for (int j = 0; j < 3; ++j) { s_idx[tid][j] = (int)idx_in[o + j]; s_w[tid][j] = __ldg(w_in + o + j); }
    }
  }
  const float* p2 = xyz2 + (size_t)b * S * 3;
  for (int pr = tid; !idx_in && pr < (S + 1) / 2; pr += QPB) {
    const int i0 = 2 * pr, i1 = 2 * pr + 1;
    const float x0 = __ldg(p2 + i0 * 3), y0 = __ldg(p2 + i0 * 3 + 1), z0 = __ldg(p2 + i0 * 3 + 2);
    float x1 = 0.f, y1 = 0.f, z1 = 0.f, n1 = __int_as_float(0x7f800000);   // odd S: the pad never enters the top 3
    if (i1 < S) { x1 = __ldg(p2 + i1 * 3); y1 = __ldg(p2 + i1 * 3 + 1); z1 = __ldg(p2 + i1 * 3 + 2); n1 = p2c_norm2_rn(x1, y1, z1); }
    sA[pr] = make_float4(x0, x1, y0, y1);
    sB[pr] = make_float4(z0, z1, p2c_norm2_rn(x0, y0, z0), n1);
  }
  __syncthreads();

  const int q = blockIdx.x * QPB + tid;
  if (q < N && !idx_in) {
    const float* p1 = xyz1 + ((size_t)b * N + q) * 3;
    const float ax = __ldg(p1), ay = __ldg(p1 + 1), az = __ldg(p1 + 2);
    const float na = p2c_norm2_rn(ax, ay, az);
    float d0 = __int_as_float(0x7f800000), d1 = d0, d2 = d0;  // +inf
    int i0 = 0, i1 = 0, i2 = 0;
    const float2 ax2 = make_float2(ax, ax), ay2 = make_float2(ay, ay), az2 = make_float2(az, az);
    const float2 na2 = make_float2(na, na), m2 = make_float2(-2.0f, -2.0f);
    auto insert = [&](float d, int j) {
      if (d < d1) {
        d2 = d1; i2 = i1;
        if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
        else { d1 = d; i1 = j; }
      } else { d2 = d; i2 = j; }
    };
#pragma unroll 4
    for (int pr = 0; pr < (S + 1) / 2; ++pr) {
      const float4 A = sA[pr], Bv = sB[pr];
      // ((-2*dot) + |q|^2) + |p|^2 with dot = fma(az,bz, fma(ay,by, ax*bx)): p2c_sqdist_expanded on both halves
      const float2 dot = __ffma2_rn(az2, make_float2(Bv.x, Bv.y),
                                    __ffma2_rn(ay2, make_float2(A.z, A.w), __fmul2_rn(ax2, make_float2(A.x, A.y))));
      const float2 d = __fadd2_rn(__fadd2_rn(__fmul2_rn(m2, dot), na2), make_float2(Bv.z, Bv.w));
      if (d.x < d2) insert(d.x, 2 * pr);
      if (d.y < d2) insert(d.y, 2 * pr + 1);
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
  if (!feats2) return;                               // search only

  const int lane = tid & 31, warp = tid >> 5;
  const float* f2 = feats2 + (size_t)b * S * ldf;
  const bool vec = (D % 4 == 0) && (ldf % 4 == 0) && (ldo % 4 == 0) &&
                   (((reinterpret_cast<uintptr_t>(feats2) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
  if (vec) {
    // two queries per iteration, their six row loads issued before either is consumed: a warp walks its 32 queries in
    // program order, so with one query per iteration exactly one L2 round trip was in flight per warp
    for (int t = warp; t < QPB; t += 2 * (QPB / 32)) {
      const int t2 = t + QPB / 32;
      const int qa = blockIdx.x * QPB + t, qb = blockIdx.x * QPB + t2;
      if (qa >= N) break;
      const bool two = qb < N;
      const int tb = two ? t2 : t;
      const float* a0 = f2 + (size_t)s_idx[t][0] * ldf;
      const float* a1 = f2 + (size_t)s_idx[t][1] * ldf;
      const float* a2 = f2 + (size_t)s_idx[t][2] * ldf;
      const float* b0 = f2 + (size_t)s_idx[tb][0] * ldf;
      const float* b1 = f2 + (size_t)s_idx[tb][1] * ldf;
      const float* b2 = f2 + (size_t)s_idx[tb][2] * ldf;
      const float w0 = s_w[t][0], w1 = s_w[t][1], w2 = s_w[t][2];
      const float x0 = s_w[tb][0], x1 = s_w[tb][1], x2 = s_w[tb][2];
      float* oa = out + ((size_t)b * N + qa) * ldo;
      float* ob = out + ((size_t)b * N + qb) * ldo;
      for (int c = lane * 4; c < D; c += 128) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(a0 + c));
        const float4 v = __ldg(reinterpret_cast<const float4*>(a1 + c));
        const float4 w = __ldg(reinterpret_cast<const float4*>(a2 + c));
        const float4 p = __ldg(reinterpret_cast<const float4*>(b0 + c));
        const float4 q = __ldg(reinterpret_cast<const float4*>(b1 + c));
        const float4 r_ = __ldg(reinterpret_cast<const float4*>(b2 + c));
        float4 r;
        r.x = __fadd_rn(__fadd_rn(__fmul_rn(u.x, w0), __fmul_rn(v.x, w1)), __fmul_rn(w.x, w2));
        r.y = __fadd_rn(__fadd_rn(__fmul_rn(u.y, w0), __fmul_rn(v.y, w1)), __fmul_rn(w.y, w2));
        r.z = __fadd_rn(__fadd_rn(__fmul_rn(u.z, w0), __fmul_rn(v.z, w1)), __fmul_rn(w.z, w2));
        r.w = __fadd_rn(__fadd_rn(__fmul_rn(u.w, w0), __fmul_rn(v.w, w1)), __fmul_rn(w.w, w2));
        *reinterpret_cast<float4*>(oa + c) = r;
        if (two) {
          r.x = __fadd_rn(__fadd_rn(__fmul_rn(p.x, x0), __fmul_rn(q.x, x1)), __fmul_rn(r_.x, x2));
          r.y = __fadd_rn(__fadd_rn(__fmul_rn(p.y, x0), __fmul_rn(q.y, x1)), __fmul_rn(r_.y, x2));
          r.z = __fadd_rn(__fadd_rn(__fmul_rn(p.z, x0), __fmul_rn(q.z, x1)), __fmul_rn(r_.z, x2));
          r.w = __fadd_rn(__fadd_rn(__fmul_rn(p.w, x0), __fmul_rn(q.w, x1)), __fmul_rn(r_.w, x2));
          *reinterpret_cast<float4*>(ob + c) = r;
        }
      }
    }
    return;
  }
  for (int t = warp; t < QPB; t += QPB / 32) {
    const int qq = blockIdx.x * QPB + t;
    if (qq >= N) break;
    const float* a0 = f2 + (size_t)s_idx[t][0] * ldf;
    const float* a1 = f2 + (size_t)s_idx[t][1] * ldf;
    const float* a2 = f2 + (size_t)s_idx[t][2] * ldf;
    const float w0 = s_w[t][0], w1 = s_w[t][1], w2 = s_w[t][2];
    float* o = out + ((size_t)b * N + qq) * ldo;
    for (int c = lane; c < D; c += 32)
      o[c] = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(a0 + c), w0), __fmul_rn(__ldg(a1 + c), w1)),
                       __fmul_rn(__ldg(a2 + c), w2));
  }
}

// Gather-only interpolation (neighbours / weights found earlier by the search half): one WARP per query, eight queries
// per iteration, for SMALL query counts (the coarse level: B x 512 queries) - the fused kernel above walks 32 queries
// per warp one after the other, which left fp2's gather (16,384 queries) at 3 warps per SM and 23 us of pure latency
// (15 us here); at 262,144 queries the staged kernel is the faster one and keeps the job.
// Same rounding as above: (u w0 + v w1) + x w2, every operation rounded.
__global__ void __launch_bounds__(256)
three_nn_gather_kernel(const float* __restrict__ feats2, int64_t ldf, const int64_t* __restrict__ idx,
                       const float* __restrict__ w, int64_t total, int N, int S, int D, float* __restrict__ out,
                       int64_t ldo, int vec) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * 8;
  // two queries per iteration (q and q + nw): both dependent chains (indices -> rows -> store) are in flight together
  for (int64_t q = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); q < total; q += 2 * nw) {
    const int64_t q2 = q + nw;
    const bool two = q2 < total;
    const int64_t qq[2] = {q, two ? q2 : q};
    const float* a[2][3];
    float wt[2][3];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const float* f2 = feats2 + (size_t)(qq[u] / N) * S * ldf;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        a[u][j] = f2 + (size_t)__ldg(idx + qq[u] * 3 + j) * ldf;
        wt[u][j] = __ldg(w + qq[u] * 3 + j);
      }
    }
    if (vec) {
      for (int c = lane * 4; c < D; c += 128) {
        float4 r[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float4 x0 = __ldg(reinterpret_cast<const float4*>(a[u][0] + c));
          const float4 x1 = __ldg(reinterpret_cast<const float4*>(a[u][1] + c));
          const float4 x2 = __ldg(reinterpret_cast<const float4*>(a[u][2] + c));
          r[u].x = __fadd_rn(__fadd_rn(__fmul_rn(x0.x, wt[u][0]), __fmul_rn(x1.x, wt[u][1])), __fmul_rn(x2.x, wt[u][2]));
          r[u].y = __fadd_rn(__fadd_rn(__fmul_rn(x0.y, wt[u][0]), __fmul_rn(x1.y, wt[u][1])), __fmul_rn(x2.y, wt[u][2]));
          r[u].z = __fadd_rn(__fadd_rn(__fmul_rn(x0.z, wt[u][0]), __fmul_rn(x1.z, wt[u][1])), __fmul_rn(x2.z, wt[u][2]));
          r[u].w = __fadd_rn(__fadd_rn(__fmul_rn(x0.w, wt[u][0]), __fmul_rn(x1.w, wt[u][1])), __fmul_rn(x2.w, wt[u][2]));
        }
        *reinterpret_cast<float4*>(out + (size_t)q * ldo + c) = r[0];
        if (two) *reinterpret_cast<float4*>(out + (size_t)q2 * ldo + c) = r[1];
      }
    } else {
      for (int c = lane; c < D; c += 32) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u == 1 && !two) break;
          out[(size_t)qq[u] * ldo + c] = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(a[u][0] + c), wt[u][0]), __fmul_rn(__ldg(a[u][1] + c), wt[u][1])),
                                                   __fmul_rn(__ldg(a[u][2] + c), wt[u][2]));
        }
      }
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
  const size_t smem = (size_t)((S + 1) / 2) * 2 * sizeof(float4);
  if (smem > 200 * 1024) return P2C_EUNSUPPORTED;
  if (smem > 48 * 1024)
    P2C_CUDA_TRY(cudaFuncSetAttribute(three_nn_interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(p2c_ceil_div(N, QPB), B);
  three_nn_interp_kernel<<<grid, QPB, smem, st>>>(xyz1, xyz2, feats2, ldf, N, S, D, out, ldo, idx_out, w_out, nullptr,
                                                   nullptr);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_three_nn_search(const float* xyz1, const float* xyz2, int B, int N, int S, int64_t* idx_out,
                                   float* w_out, void* stream) {
  if (!xyz1 || !xyz2 || !idx_out || !w_out || B <= 0 || N <= 0 || S <= 0) return P2C_EINVAL;
  if (S < 3) return P2C_EUNSUPPORTED;
  const size_t smem = (size_t)((S + 1) / 2) * 2 * sizeof(float4);
  if (smem > 200 * 1024) return P2C_EUNSUPPORTED;
  if (smem > 48 * 1024)
    P2C_CUDA_TRY(cudaFuncSetAttribute(three_nn_interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(p2c_ceil_div(N, QPB), B);
  three_nn_interp_kernel<<<grid, QPB, smem, (cudaStream_t)stream>>>(xyz1, xyz2, nullptr, 0, N, S, 0, nullptr, 0, idx_out,
                                                                    w_out, nullptr, nullptr);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_three_nn_gather(const float* feats2, int64_t ldf, const int64_t* idx, const float* w, int B, int N,
                                   int S, int D, float* out, int64_t ldo, void* stream) {
  if (!feats2 || !idx || !w || !out || B <= 0 || N <= 0 || S <= 0 || D <= 0 || ldf < D || ldo < D) return P2C_EINVAL;
  const int64_t total = (int64_t)B * N;
  if (total > 65536) {
    // many queries: the staged kernel (neighbours / weights of 128 queries through shared memory, 32 queries per warp)
    // runs at the L2 bandwidth - measured 55 us against 82 us for the warp-per-query kernel at 262,144 queries
    dim3 grid(p2c_ceil_div(N, QPB), B);
    three_nn_interp_kernel<<<grid, QPB, 0, (cudaStream_t)stream>>>(nullptr, nullptr, feats2, ldf, N, S, D, out, ldo,
                                                                   nullptr, nullptr, idx, w);
    P2C_RETURN_IF_CUDA_ERROR();
    return 0;
  }
  const int vec = (D % 4 == 0) && (ldf % 4 == 0) && (ldo % 4 == 0) &&
                  (((reinterpret_cast<uintptr_t>(feats2) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
  const int blocks = (int)min((int64_t)148 * 8, (total + 15) / 16);
  three_nn_gather_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(feats2, ldf, idx, w, total, N, S, D, out, ldo, vec);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
