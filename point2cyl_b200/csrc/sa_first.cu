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
#include "common.cuh"

namespace {

template <int CPL>   // channels per lane: 2 (C = 64) or 4 (C = 128)
__global__ void __launch_bounds__(256)
sa_first_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, const int64_t* __restrict__ idx,
                const float* __restrict__ Qf, int64_t ldq, const float* __restrict__ W, int64_t ldw,
                const float* __restrict__ bias, int N, int S, int ns, int64_t rows, float* __restrict__ Y,
                int64_t ldy, double* __restrict__ stats) {
  constexpr int C = 32 * CPL;
  __shared__ float s_red[8][2][C];
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
  float s1[CPL], s2[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) { s1[i] = 0.f; s2[i] = 0.f; }

  // UNR rows per warp and step: the index -> coordinates / Qf row -> store chain is two dependent L2 round
  // trips, so independent rows are kept in flight
  constexpr int UNR = 4;
  const int64_t wstride = (int64_t)gridDim.x * 8 * UNR;
  for (int64_t r0 = ((int64_t)blockIdx.x * 8 + warp) * UNR; r0 < rows; r0 += wstride) {
    int64_t p[UNR], bs[UNR], bb[UNR];
    bool ok[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      ok[u] = r0 + u < rows;
      const int64_t r = ok[u] ? r0 + u : rows - 1;
      bs[u] = r / ns;
      bb[u] = bs[u] / S;
      p[u] = __ldg(idx + r);
    }
    float d[UNR][3];
    float y[UNR][CPL];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      p[u] = (p[u] < 0 || p[u] >= N) ? 0 : p[u];
      const float* pp = xyz + ((size_t)bb[u] * N + p[u]) * 3;
      const float* cc = new_xyz + (size_t)bs[u] * 3;
      d[u][0] = __ldg(pp) - __ldg(cc); d[u][1] = __ldg(pp + 1) - __ldg(cc + 1); d[u][2] = __ldg(pp + 2) - __ldg(cc + 2);
      if (Qf) {
        const float* q = Qf + ((size_t)bb[u] * N + p[u]) * ldq + c0;
        if (CPL == 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(q));
          y[u][0] = t.x; y[u][1] = t.y; y[u][2] = t.z; y[u][CPL - 1] = t.w;
        } else {
          const float2 t = __ldg(reinterpret_cast<const float2*>(q));
          y[u][0] = t.x; y[u][1] = t.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < CPL; ++i) y[u][i] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (!ok[u]) continue;
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        y[u][i] += fmaf(wx[i][2], d[u][2], fmaf(wx[i][1], d[u][1], wx[i][0] * d[u][0])) + bj[i];
        s1[i] += y[u][i];
        s2[i] = fmaf(y[u][i], y[u][i], s2[i]);
      }
      float* yo = Y + (size_t)(r0 + u) * ldy + c0;
      if (CPL == 4) *reinterpret_cast<float4*>(yo) = make_float4(y[u][0], y[u][1], y[u][2], y[u][CPL - 1]);
      else *reinterpret_cast<float2*>(yo) = make_float2(y[u][0], y[u][1]);
    }
  }
  if (stats) {
#pragma unroll
    for (int i = 0; i < CPL; ++i) { s_red[warp][0][c0 + i] = s1[i]; s_red[warp][1][c0 + i] = s2[i]; }
    __syncthreads();
    for (int e = threadIdx.x; e < 2 * C; e += 256) {
      const int which = e / C, c = e - which * C;
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += (double)s_red[w][which][c];
      atomicAdd(stats + which * C + c, t);
    }
  }
}

}  // namespace

extern "C" int p2c_sa_first_layer(const float* xyz, const float* new_xyz, const int64_t* idx, const float* Qf,
                                  int64_t ldq, const float* W, int64_t ldw, const float* bias, int B, int N, int S,
                                  int nsample, int C, float* Y, int64_t ldy, double* stats, void* stream) {
  if (!xyz || !new_xyz || !idx || !W || !Y || B <= 0 || N <= 0 || S <= 0 || nsample <= 0 || ldw < 3 || ldy < C)
    return P2C_EINVAL;
  if (C != 64 && C != 128) return P2C_EUNSUPPORTED;
  if ((ldy % 4) != 0 || (reinterpret_cast<uintptr_t>(Y) & 15) != 0) return P2C_EALIGN;
  if (Qf && ((ldq % 4) != 0 || (reinterpret_cast<uintptr_t>(Qf) & 15) != 0 || ldq < C)) return P2C_EALIGN;
  const int64_t rows = (int64_t)B * S * nsample;
  const int blocks = (int)min((int64_t)148 * 8, (rows + 31) / 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 128)
    sa_first_kernel<4><<<blocks, 256, 0, st>>>(xyz, new_xyz, idx, Qf, ldq, W, ldw, bias, N, S, nsample, rows, Y, ldy, stats);
  else
    sa_first_kernel<2><<<blocks, 256, 0, st>>>(xyz, new_xyz, idx, Qf, ldq, W, ldw, bias, N, S, nsample, rows, Y, ldy, stats);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
