// First MLP layer of a set-abstraction level, fused with the grouping gather (replaces index_points + centring +
// concat + the first Conv2d of models/pointnet_util.py:130-139,200-201):
//
//   Y[r, c] = sum_d Wx[c, d] * (xyz[b, p, d] - new_xyz[b, s, d])  +  Qf[b*N + p, c]  +  bias[c]
//   r = (b, s, j),  p = idx[b, s, j]
//
// The 1x1 convolution is linear, so the feature half of the layer commutes with the gather:
// W_f * feats[p] = (W_f * feats)[p] = Qf[p], computed ONCE PER SOURCE POINT (N rows) by p2c_linear instead of
// once per (centre, neighbour) pair (S*nsample rows, 64x more).  The grouped input tensor (B,S,ns,3+D) - 138 MB
// at level 2 - is never written or read; the xyz half is three FMAs per output on the exactly-centred
// coordinates.  The kernel is HBM bound on its only large stream, the raw output Y (4*C bytes per row);
// Qf rows come from L2.  BatchNorm sum / sum-of-squares are accumulated per lane and folded once per CTA.
// One warp per row, lane = C/32 consecutive channels.
#include <cstdlib>
#include "bn_fold.cuh"

namespace {

template <int CPL>   // channels per lane: 2 (C = 64) or 4 (C = 128)
__global__ void __launch_bounds__(256)
sa_first_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, const int64_t* __restrict__ idx,
                const float* __restrict__ Qf, int64_t ldq, const float* __restrict__ W, int64_t ldw,
                const float* __restrict__ bias, int N, int S, int ns, int64_t rows, float* __restrict__ Y,
                int64_t ldy, double* __restrict__ stats) {
  constexpr int C = 32 * CPL;
  __shared__ double s_red[8][2][C];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = lane * CPL;
  float wx[CPL][3], bj[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    wx[i][0] = __ldg(W + (size_t)(c0 + i) * ldw + 0);
    wx[i][1] = __ldg(W + (size_t)(c0 + i) * ldw + 1);
    wx[i][2] = __ldg(W + (size_t)(c0 + i) * ldw + 2);
    bj[i] = bias ? __ldg(bias + c0 + i) : 0.f;
  }
  // BatchNorm sums: fp32 within a 32-row step, fp64 across steps (var = E[y^2] - mean^2 cancels, so the sums must
  // carry more than fp32 once a lane has seen hundreds of rows)
  double s1[CPL], s2[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) { s1[i] = 0.0; s2[i] = 0.0; }

  // 32 rows per warp and step: lane l fetches the index and the centred coordinates of row r0 + l (one coalesced
  // 256-byte index load and three gathers for 32 rows instead of a dependent L2 round-trip chain per row), then the
  // warp walks the 32 rows with the coordinates broadcast by shuffles; lane = CPL consecutive output channels.
  const int64_t wstride = (int64_t)gridDim.x * 8 * 32;
  for (int64_t r0 = ((int64_t)blockIdx.x * 8 + warp) * 32; r0 < rows; r0 += wstride) {
    const int64_t rl = r0 + lane;
    const bool okl = rl < rows;
    const unsigned rr = (unsigned)(okl ? rl : rows - 1);             // rows < 2^31 (checked by the entry point)
    const unsigned bsl = rr / (unsigned)ns, bl = bsl / (unsigned)S;
    int64_t pl = __ldg(idx + rr);
    pl = (pl < 0 || pl >= N) ? 0 : pl;
    const unsigned src = bl * (unsigned)N + (unsigned)pl;            // source row of xyz / Qf (B*N < 2^32)
    const float* pp = xyz + (size_t)src * 3;
    const float* cc = new_xyz + (size_t)bsl * 3;
    const float dl0 = __ldg(pp) - __ldg(cc), dl1 = __ldg(pp + 1) - __ldg(cc + 1), dl2 = __ldg(pp + 2) - __ldg(cc + 2);
    const int nrow = (int)min((int64_t)32, rows - r0);
    float t1[CPL], t2[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) { t1[i] = 0.f; t2[i] = 0.f; }
    // eight rows at a time, their feature gathers issued before any of them is consumed: the kernel is bound by the
    // latency of those loads (ncu: issue slots 30 % busy, long-scoreboard stalls), not by their bandwidth
    for (int j0 = 0; j0 < nrow; j0 += 8) {
      float y[8][CPL];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) y[u][i] = 0.f;
        if (Qf && j0 + u < nrow) {
          const unsigned sj = __shfl_sync(P2C_FULL_MASK, src, j0 + u);
          const float* q = Qf + (size_t)sj * ldq + c0;
          if (CPL == 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(q));
            y[u][0] = t.x; y[u][1] = t.y; y[u][2] = t.z; y[u][CPL - 1] = t.w;
          } else {
            const float2 t = __ldg(reinterpret_cast<const float2*>(q));
            y[u][0] = t.x; y[u][1] = t.y;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u;
        if (j >= nrow) break;
        const float d0 = __shfl_sync(P2C_FULL_MASK, dl0, j), d1 = __shfl_sync(P2C_FULL_MASK, dl1, j),
                    d2 = __shfl_sync(P2C_FULL_MASK, dl2, j);
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          y[u][i] += fmaf(wx[i][2], d2, fmaf(wx[i][1], d1, wx[i][0] * d0)) + bj[i];
          t1[i] += y[u][i];
          t2[i] = fmaf(y[u][i], y[u][i], t2[i]);
        }
        float* yo = Y + (size_t)(r0 + j) * ldy + c0;
        if (CPL == 4) *reinterpret_cast<float4*>(yo) = make_float4(y[u][0], y[u][1], y[u][2], y[u][CPL - 1]);
        else *reinterpret_cast<float2*>(yo) = make_float2(y[u][0], y[u][1]);
      }
    }
#pragma unroll
    for (int i = 0; i < CPL; ++i) { s1[i] += (double)t1[i]; s2[i] += (double)t2[i]; }
  }
  if (stats) {
#pragma unroll
    for (int i = 0; i < CPL; ++i) { s_red[warp][0][c0 + i] = s1[i]; s_red[warp][1][c0 + i] = s2[i]; }
    __syncthreads();
    for (int e = threadIdx.x; e < 2 * C; e += 256) {
      const int which = e / C, c = e - which * C;
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s_red[w][which][c];
      atomicAdd(stats + which * C + c, t);
    }
  }
}

// ---- the first layer WITHOUT its output (p2c_sa_xyz_linear): BatchNorm statistics in closed form ----------------
// With no point features the first conv sees only the centred neighbour coordinates d = xyz[idx] - centre, so its
// raw output y_c = w_c . d + b_c never has to exist: the next layer recomputes it in its operand transform
// (linear_tc.cu), and the train-mode BatchNorm statistics of y follow from nine moments of d over all rows:
//   sum y_c   = n b_c + w_c . S1                      S1 = sum d      (3)
//   sum y_c^2 = w_c^T S2 w_c + 2 b_c w_c . S1 + n b_c^2   S2 = sum d d^T  (6)
// group_moments_kernel reduces (S1, S2) per CTA (coordinates only: it belongs to the geometry stage),
// sa_xyz_stats_kernel sums the CTA partials and writes the 2 C sums where p2c_sa_first_layer would have accumulated them.
constexpr int MOM_CTAS = 592;     // 4 per SM: the per-row work is a dependent index -> coordinate load chain

__global__ void __launch_bounds__(256)
group_moments_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, const int64_t* __restrict__ idx,
                     int N, int S, int ns, int64_t rows, double* __restrict__ partials) {
  __shared__ double s_red[8][9];
  double m[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) m[q] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t r0 = (int64_t)blockIdx.x * 256 + threadIdx.x; r0 < rows; r0 += 4 * stride) {
    // four independent rows per step: the loads of all four chains are in flight together
    float d[4][3];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = r0 + u * stride;
      d[u][0] = d[u][1] = d[u][2] = 0.f;
      if (r < rows) {
        const unsigned bs = (unsigned)r / (unsigned)ns, b = bs / (unsigned)S;
        int64_t p = __ldg(idx + r);
        p = (p < 0 || p >= N) ? 0 : p;
        const float* pp = xyz + ((size_t)b * N + (size_t)p) * 3;
        const float* cc = new_xyz + (size_t)bs * 3;
        d[u][0] = __ldg(pp) - __ldg(cc); d[u][1] = __ldg(pp + 1) - __ldg(cc + 1); d[u][2] = __ldg(pp + 2) - __ldg(cc + 2);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double d0 = (double)d[u][0], d1 = (double)d[u][1], d2 = (double)d[u][2];
      m[0] += d0; m[1] += d1; m[2] += d2;
      m[3] += d0 * d0; m[4] += d0 * d1; m[5] += d0 * d2; m[6] += d1 * d1; m[7] += d1 * d2; m[8] += d2 * d2;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 9; ++q) {
    const double v = p2c_warp_sum(m[q]);
    if (lane == 0) s_red[warp][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 9 + threadIdx.x] = t;
  }
  // the last CTA to finish sums the partials in a fixed order (deterministic) into partials[9 P .. 9 P + 9) and
  // re-arms the launch counter (the word after them) for the next launch
  __shared__ bool s_last;
  unsigned* counter = reinterpret_cast<unsigned*>(partials + (size_t)gridDim.x * 9 + 9);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
#pragma unroll
  for (int q = 0; q < 9; ++q) m[q] = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) {
#pragma unroll
    for (int q = 0; q < 9; ++q) m[q] += __ldcg(partials + (size_t)i * 9 + q);
  }
#pragma unroll
  for (int q = 0; q < 9; ++q) {
    const double v = p2c_warp_sum(m[q]);
    if (lane == 0) s_red[warp][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[w][threadIdx.x];
    partials[(size_t)gridDim.x * 9 + threadIdx.x] = t;
  }
  if (threadIdx.x == 0) *counter = 0u;
}

__global__ void __launch_bounds__(128)
sa_xyz_stats_kernel(const double* __restrict__ moments, double count, const float* __restrict__ W, int64_t ldw,
                    const float* __restrict__ bias, int C, double* __restrict__ stats) {
  for (int c = threadIdx.x; c < C; c += 128) {
    double sum, sumsq;
    p2c_xyz_first_sums(moments, count, (double)__ldg(W + (size_t)c * ldw), (double)__ldg(W + (size_t)c * ldw + 1),
                       (double)__ldg(W + (size_t)c * ldw + 2), bias ? (double)__ldg(bias + c) : 0.0, sum, sumsq);
    stats[c] = sum;
    stats[C + c] = sumsq;
  }
}

}  // namespace

// linear_tc.cu
int p2c_linear_tc(const float* X, int64_t ldx, const float* W, const float* bias, const float* in_scale,
                  const float* in_shift, const float* in_mask, int64_t ldmask, float* Y, int64_t ldy, int M, int N,
                  int K, double* stats, int pool_group, float* Ymax, float* Ymin, int precision,
                  const p2c_bn_fold* in_bn, cudaStream_t st, const int64_t* drop_seed, const P2cXyzFirst* xyz_first);

extern "C" int p2c_group_moments_size(void) { return MOM_CTAS * 9 + 9 + 1; }

extern "C" int p2c_group_moments(const float* xyz, const float* new_xyz, const int64_t* idx, int B, int N, int S,
                                 int nsample, double* partials, void* stream) {
  if (!xyz || !new_xyz || !idx || !partials || B <= 0 || N <= 0 || S <= 0 || nsample <= 0) return P2C_EINVAL;
  const int64_t rows = (int64_t)B * S * nsample;
  if (rows >= ((int64_t)1 << 31) || (int64_t)B * N >= ((int64_t)1 << 32)) return P2C_EUNSUPPORTED;
  group_moments_kernel<<<MOM_CTAS, 256, 0, (cudaStream_t)stream>>>(xyz, new_xyz, idx, N, S, nsample, rows, partials);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_sa_xyz_stats(const double* partials, int64_t rows, const float* W, int64_t ldw, const float* bias,
                                int C, double* stats, void* stream) {
  if (!partials || !W || !stats || rows <= 0 || C <= 0 || ldw < 3) return P2C_EINVAL;
  sa_xyz_stats_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(partials + MOM_CTAS * 9, (double)rows, W, ldw, bias, C, stats);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

int p2c_sa_pair_tc(const P2cXyzFirst& g, int B, const float* scale0, const float* shift0, const p2c_bn_fold* bn0,
                   const float* W1, const float* b1, int C0, int N1, float* Y, int64_t ldy, double* stats,
                   cudaStream_t st);
static bool getenv_pair_enabled() {            // P2C_SA_PAIR=0: A/B switch back to the channels-as-lanes kernel
  static int on = -1;
  if (on < 0) { const char* e = getenv("P2C_SA_PAIR"); on = (e && e[0] == '0') ? 0 : 1; }
  return on != 0;
}

extern "C" int p2c_sa_xyz_linear(const float* xyz, const float* new_xyz, const int64_t* idx, int B, int N, int S,
                                 int nsample, const float* W0, int64_t ldw0, const float* b0, int C0,
                                 const float* scale0, const float* shift0, const p2c_bn_fold* bn0,
                                 const double* moments, const float* W1,
                                 const float* b1, int N1, float* Y, int64_t ldy, double* stats, int pool_group,
                                 float* Ymax, float* Ymin, void* stream) {
  if (!xyz || !new_xyz || !idx || !W0 || !W1 || B <= 0 || N <= 0 || S <= 0 || nsample <= 0 || C0 <= 0 || N1 <= 0 ||
      ldw0 < 3)
    return P2C_EINVAL;
  if ((scale0 == nullptr) != (shift0 == nullptr)) return P2C_EINVAL;
  if (int e = p2c_bn_fold_check(bn0, C0)) return e;
  if (!Y && !pool_group && !stats) return P2C_EINVAL;
  if (Y && ldy < N1) return P2C_EINVAL;
  const int64_t rows = (int64_t)B * S * nsample;
  if (rows >= ((int64_t)1 << 31) || (int64_t)B * N >= ((int64_t)1 << 32)) return P2C_EUNSUPPORTED;
  if (pool_group && (!Ymax || !Ymin || rows % pool_group != 0)) return P2C_EINVAL;
  if (moments && bn0 && bn0->stats && bn0->count != rows) return P2C_EINVAL;
  const P2cXyzFirst g{xyz, new_xyz, idx, W0, ldw0, b0, N, S, nsample, moments ? moments + MOM_CTAS * 9 : nullptr};
  if (!pool_group && Y && getenv_pair_enabled()) {
    // a 64 -> 64 second layer whose rows are kept: rows-as-lanes form (half the tensor time), sa_stack_tc.cu MODE 1
    const int rc = p2c_sa_pair_tc(g, B, scale0, shift0, bn0, W1, b1, C0, N1, Y, ldy, stats, (cudaStream_t)stream);
    if (rc != P2C_EUNSUPPORTED) return rc;
  }
  return p2c_linear_tc(nullptr, 0, W1, b1, scale0, shift0, nullptr, 0, Y, ldy, (int)rows, N1, C0, stats, pool_group,
                       Ymax, Ymin, P2C_PREC_3XTF32, bn0, (cudaStream_t)stream, nullptr, &g);
}

extern "C" int p2c_sa_first_layer(const float* xyz, const float* new_xyz, const int64_t* idx, const float* Qf,
                                  int64_t ldq, const float* W, int64_t ldw, const float* bias, int B, int N, int S,
                                  int nsample, int C, float* Y, int64_t ldy, double* stats, void* stream) {
  if (!xyz || !new_xyz || !idx || !W || !Y || B <= 0 || N <= 0 || S <= 0 || nsample <= 0 || ldw < 3 || ldy < C)
    return P2C_EINVAL;
  if (C != 64 && C != 128) return P2C_EUNSUPPORTED;
  if ((ldy % 4) != 0 || (reinterpret_cast<uintptr_t>(Y) & 15) != 0) return P2C_EALIGN;
  if (Qf && ((ldq % 4) != 0 || (reinterpret_cast<uintptr_t>(Qf) & 15) != 0 || ldq < C)) return P2C_EALIGN;
  const int64_t rows = (int64_t)B * S * nsample;
  if (rows >= ((int64_t)1 << 31) || (int64_t)B * N >= ((int64_t)1 << 32)) return P2C_EUNSUPPORTED;
  const int blocks = (int)min((int64_t)148 * 8, (rows + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 128)
    sa_first_kernel<4><<<blocks, 256, 0, st>>>(xyz, new_xyz, idx, Qf, ldq, W, ldw, bias, N, S, nsample, rows, Y, ldy, stats);
  else
    sa_first_kernel<2><<<blocks, 256, 0, st>>>(xyz, new_xyz, idx, Qf, ldq, W, ldw, bias, N, S, nsample, rows, Y, ldy, stats);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
