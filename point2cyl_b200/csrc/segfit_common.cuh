// Shared pieces of the loss-side kernels (segfit.cu forward, segfit_bwd.cu backward): the per-cloud statistics
// layout, the KP-lanes-per-point helpers and the 3x3 Jacobi eigen-solve.
#pragma once
#include "common.cuh"

namespace {

constexpr int SEG_THREADS = 256;
constexpr int SEG_CHUNK = 256;   // points per CTA (B=32 x N=8192: 1024 CTAs; with 1024 points per CTA the 256-CTA grid was latency bound)

__host__ __device__ inline int seg_stride(int K) { return K * K + 19 * K + 2; }
__host__ __device__ inline int off_cnt(int K) { return K * K; }
__host__ __device__ inline int off_cbar(int K) { return K * K + K; }
__host__ __device__ inline int off_cbase(int K) { return K * K + 2 * K; }
__host__ __device__ inline int off_colsum(int K) { return K * K + 3 * K; }
__host__ __device__ inline int off_C(int K) { return K * K + 4 * K; }
__host__ __device__ inline int off_Mbar(int K) { return K * K + 7 * K; }
__host__ __device__ inline int off_Mbase(int K) { return K * K + 13 * K; }
__host__ __device__ inline int off_normal(int K) { return K * K + 19 * K; }
__host__ __device__ inline int off_maxlab(int K) { return K * K + 19 * K + 1; }

template <int KP>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = KP / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(P2C_FULL_MASK, v, o));
  return v;
}
template <int KP>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = KP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(P2C_FULL_MASK, v, o);
  return v;
}
// sum over the lanes of a warp that own the same column (stride KP)
template <int KP>
__device__ __forceinline__ float column_sum(float v) {
#pragma unroll
  for (int o = 16; o >= KP; o >>= 1) v += __shfl_xor_sync(P2C_FULL_MASK, v, o);
  return v;
}

// softmax over the 2K logits of one point, spread over the KP lanes of its group
template <int KP>
__device__ __forceinline__ void point_softmax(const float* __restrict__ wrow, int k, int K,
                                              float& raw_bar, float& raw_base, float& wb, float& wc) {
  const float NEG = -__int_as_float(0x7f800000);
  raw_bar = NEG; raw_base = NEG;
  if (k < K) { raw_bar = __ldg(wrow + 2 * k); raw_base = __ldg(wrow + 2 * k + 1); }
  const float m = group_max<KP>(fmaxf(raw_bar, raw_base));
  const float eb = k < K ? expf(raw_bar - m) : 0.f;
  const float ec = k < K ? expf(raw_base - m) : 0.f;
  const float s = group_sum<KP>(eb + ec);
  wb = eb / s;
  wc = ec / s;
}

// ---- 3x3 symmetric eigen-solve: cyclic Jacobi in float64, eigenvalues ascending ------------------
__device__ void jacobi3(const double a_in[6] /* xx xy xz yy yz zz */, double eval[3], double evec[3][3]) {
  double a[3][3] = {{a_in[0], a_in[1], a_in[2]}, {a_in[1], a_in[3], a_in[4]}, {a_in[2], a_in[4], a_in[5]}};
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 32; ++sweep) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-32 * diag || off == 0.0) break;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0;
      const int q = pq == 0 ? 1 : 2;
      const double apq = a[p][q];
      if (apq == 0.0) continue;
      const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
      const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
      for (int r = 0; r < 3; ++r) {  // A <- A J
        const double arp = a[r][p], arq = a[r][q];
        a[r][p] = c * arp - s * arq;
        a[r][q] = s * arp + c * arq;
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {  // A <- J^T A
        const double apr = a[p][r], aqr = a[q][r];
        a[p][r] = c * apr - s * aqr;
        a[q][r] = s * apr + c * aqr;
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {  // V <- V J
        const double vrp = v[r][p], vrq = v[r][q];
        v[r][p] = c * vrp - s * vrq;
        v[r][q] = s * vrp + c * vrq;
      }
    }
  }
  int order[3] = {0, 1, 2};
  double d[3] = {a[0][0], a[1][1], a[2][2]};
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2 - i; ++j)
      if (d[order[j]] > d[order[j + 1]]) { int t = order[j]; order[j] = order[j + 1]; order[j + 1] = t; }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    eval[i] = d[order[i]];
    double x = v[0][order[i]], y = v[1][order[i]], z = v[2][order[i]];
    const double n = sqrt(x * x + y * y + z * z);
    // canonical sign: the component of largest magnitude is positive (LAPACK's sign is arbitrary)
    const double ax = fabs(x), ay = fabs(y), az = fabs(z);
    const double big = (ax >= ay && ax >= az) ? x : (ay >= az ? y : z);
    const double sgn = (big < 0.0 ? -1.0 : 1.0) / (n > 0.0 ? n : 1.0);
    evec[i][0] = x * sgn; evec[i][1] = y * sgn; evec[i][2] = z * sgn;
  }
}

inline int kp_of(int K) { return K <= 2 ? 2 : K <= 4 ? 4 : K <= 8 ? 8 : K <= 16 ? 16 : 0; }

}  // namespace
