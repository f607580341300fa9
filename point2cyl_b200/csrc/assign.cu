// On-device assignment for the segmentation loss (replaces the scipy call at losses.py:43).
//
// Per cloud: maximise sum_g score[g, match[g]] over injective maps of the n_gt ground-truth rows into the K
// predicted columns (n_gt <= K <= 16).  Classic Hungarian algorithm with row/column potentials (shortest augmenting
// paths), float64 like scipy's linear_sum_assignment.  One WARP per cloud, lane j = column j (lane 0 is the
// algorithm's virtual column 0): the two inner loops over the columns - relax the reduced costs, pick the minimum,
// update the potentials - are lane-parallel with warp shuffles, everything lives in registers (no local-memory
// arrays), so the n * K sequential steps cost a few dozen cycles each instead of a K-long serial loop.  It returns
// an exact optimum; when several optima tie exactly the choice may differ from scipy's (cannot happen for generic
// real scores).  Slots >= n_gt keep match 0 like the reference (losses.py:30,45).
#include "common.cuh"

namespace {

constexpr int KMAX = 16;

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(P2C_FULL_MASK, v, src); }

__global__ void __launch_bounds__(128)
hungarian_kernel(const float* __restrict__ score, const int32_t* __restrict__ n_gt, int B, int K,
                 int64_t* __restrict__ match) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* s = score + (size_t)b * K * K;
  int n = n_gt[b];
  n = n < 0 ? 0 : (n > K ? K : n);
  // lane j (1..K) owns column j: potential v, matched row p (0 = none), back pointer way; lane i (1..n) also holds
  // the row potential u[i]
  const bool col = lane >= 1 && lane <= K;
  double v = 0.0, u = 0.0;
  int p = 0, way = 0;
  for (int i = 1; i <= n; ++i) {
    // p[0] = i, j0 = 0
    if (lane == 0) p = i;
    int j0 = 0;
    double minv = 1e300;
    bool used = false;
    while (true) {
      if (lane == j0) used = true;
      const int i0 = __shfl_sync(P2C_FULL_MASK, p, j0);
      const double u_i0 = shfl_d(u, i0);
      double cur = 1e300;
      if (col && !used) {
        cur = -(double)__ldg(s + (i0 - 1) * K + (lane - 1)) - u_i0 - v;
        if (cur < minv) { minv = cur; way = j0; }
      }
      // delta = min over unused columns of minv, j1 = the LOWEST such column (the serial loop keeps the first minimum)
      double delta = (col && !used) ? minv : 1e300;
      int j1 = (col && !used) ? lane : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double od = shfl_d(delta, lane ^ o);
        const int oj = __shfl_xor_sync(P2C_FULL_MASK, j1, o);
        if (od < delta || (od == delta && oj < j1)) { delta = od; j1 = oj; }
      }
      // potentials: used columns (including the virtual column 0) shift, unused ones shrink their slack
      const int prow = p;                       // row matched to my column (lane 0: the row being inserted)
      const unsigned used_mask = __ballot_sync(P2C_FULL_MASK, used && (lane == 0 || col));
      if (used) v -= delta; else if (col) minv -= delta;
      // u[p[j]] += delta for every used column j: lane r (a row) adds delta once per used column matched to it
      {
        double add = 0.0;
        for (unsigned m = used_mask; m; m &= m - 1) {
          const int jj = __ffs(m) - 1;
          const int rr = __shfl_sync(P2C_FULL_MASK, prow, jj);
          if (rr == lane) add += delta;
        }
        u += add;
      }
      j0 = j1;
      const int pj = __shfl_sync(P2C_FULL_MASK, p, j0);
      if (pj == 0) break;
    }
    // augment along the way pointers
    while (j0) {
      const int j1 = __shfl_sync(P2C_FULL_MASK, way, j0);
      const int pv = __shfl_sync(P2C_FULL_MASK, p, j1);
      if (lane == j0) p = pv;
      j0 = j1;
    }
  }
  int64_t* m = match + (size_t)b * K;
  if (lane < K) m[lane] = 0;
  __syncwarp();
  if (col && p > 0) m[p - 1] = lane - 1;
}

}  // namespace

extern "C" int p2c_hungarian(const float* score, const int32_t* n_gt, int B, int K, int64_t* match, void* stream) {
  if (!score || !n_gt || !match || B <= 0 || K <= 0) return P2C_EINVAL;
  if (K > KMAX) return P2C_EUNSUPPORTED;
  hungarian_kernel<<<p2c_ceil_div(B, 4), 128, 0, (cudaStream_t)stream>>>(score, n_gt, B, K, match);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
