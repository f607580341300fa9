// On-device assignment for the segmentation loss (replaces the scipy call at losses.py:43).
//
// Per cloud: maximise sum_g score[g, match[g]] over injective maps of the n_gt ground-truth rows into the K
// predicted columns (n_gt <= K <= 16).  Classic O(n^2 m) Hungarian algorithm with row/column potentials
// (shortest augmenting paths), float64 like scipy's linear_sum_assignment, one thread per cloud - K^3 <= 4096
// steps, so the whole batch costs less than the device->host copy it replaces.  It returns an exact optimum;
// when several optima tie exactly the choice may differ from scipy's (cannot happen for generic real scores).
// Slots >= n_gt keep match 0 like the reference (losses.py:30,45).
#include "common.cuh"

namespace {

constexpr int KMAX = 16;

__global__ void hungarian_kernel(const float* __restrict__ score, const int32_t* __restrict__ n_gt, int B, int K,
                                 int64_t* __restrict__ match) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* s = score + (size_t)b * K * K;
  int n = n_gt[b];
  n = n < 0 ? 0 : (n > K ? K : n);
  // minimise cost = -score; 1-based indices, p[j] = row matched to column j (0 = none)
  double u[KMAX + 1], v[KMAX + 1], minv[KMAX + 1];
  int p[KMAX + 1], way[KMAX + 1];
  bool used[KMAX + 1];
  for (int j = 0; j <= K; ++j) { v[j] = 0.0; p[j] = 0; way[j] = 0; }
  for (int i = 0; i <= n; ++i) u[i] = 0.0;
  for (int i = 1; i <= n; ++i) {
    p[0] = i;
    int j0 = 0;
    for (int j = 0; j <= K; ++j) { minv[j] = 1e300; used[j] = false; }
    do {
      used[j0] = true;
      const int i0 = p[j0];
      double delta = 1e300;
      int j1 = 0;
      for (int j = 1; j <= K; ++j) {
        if (used[j]) continue;
        const double cur = -(double)s[(i0 - 1) * K + (j - 1)] - u[i0] - v[j];
        if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
        if (minv[j] < delta) { delta = minv[j]; j1 = j; }
      }
      for (int j = 0; j <= K; ++j) {
        if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
        else minv[j] -= delta;
      }
      j0 = j1;
    } while (p[j0] != 0);
    do {
      const int j1 = way[j0];
      p[j0] = p[j1];
      j0 = j1;
    } while (j0);
  }
  int64_t* m = match + (size_t)b * K;
  for (int g = 0; g < K; ++g) m[g] = 0;
  for (int j = 1; j <= K; ++j)
    if (p[j] > 0) m[p[j] - 1] = j - 1;
}

}  // namespace

extern "C" int p2c_hungarian(const float* score, const int32_t* n_gt, int B, int K, int64_t* match, void* stream) {
  if (!score || !n_gt || !match || B <= 0 || K <= 0) return P2C_EINVAL;
  if (K > KMAX) return P2C_EUNSUPPORTED;
  hungarian_kernel<<<p2c_ceil_div(B, 32), 32, 0, (cudaStream_t)stream>>>(score, n_gt, B, K, match);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
