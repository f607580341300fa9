// Projection / scale / extent closed forms of the extrusion fit and the eval-side helpers
// (row a18).  Replaces the K x B Python loops of data_utils.py:1014-1417 (sketch_implicit_projection{,2,3}),
// :1650-1730 (get_extrusion_extents), eval.py:409-436 (segment centroids) and the elementwise eval metrics of
// losses.py:55-68 (hard_W_encoding) / :146-159 (compute_normal_difference).
//
// The reference picks the barrel points of a segment with `nonzero()` (ascending point index), draws
// `torch.randint(0, n, (S,))` on the CPU generator per (segment, cloud) and gathers.  Here:
//   p2c_segment_lists   one CTA per (cloud, segment): order-preserving compaction of the member point indices
//                       (ballot + popcount prefix), member count per (cloud, segment);
//   p2c_sketch_project  one CTA per (segment, cloud): rotation that takes the axis to +z — Rodrigues on the
//                       UN-normalised vector (a x z)*angle exactly as the reference feeds torchgeometry — gather of
//                       the sampled members, p*R, drop z, subtract the projected centre, max-norm scale;
//   p2c_extrusion_extents  same selection, min/max of (p - c).a.
// The host keeps only what must stay there to reproduce the reference's random stream: the randint draws.
#include "common.cuh"

namespace {

constexpr int LIST_THREADS = 256;

// members of list (b,k): seg[b,n] == k and (bb == NULL or bb[b,n] == bb_value)
__global__ void __launch_bounds__(LIST_THREADS)
segment_lists_kernel(const int64_t* __restrict__ seg, const int64_t* __restrict__ bb, int bb_value, int N, int K,
                     int32_t* __restrict__ counts, int32_t* __restrict__ lists) {
  const int k = blockIdx.x, b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ int warp_tot[LIST_THREADS / 32];
  __shared__ int base_s;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  const int64_t* s = seg + (size_t)b * N;
  const int64_t* t = bb ? bb + (size_t)b * N : nullptr;
  int32_t* out = lists ? lists + ((size_t)b * K + k) * N : nullptr;
  for (int n0 = 0; n0 < N; n0 += LIST_THREADS) {
    const int n = n0 + threadIdx.x;
    bool in = false;
    if (n < N) in = (s[n] == (int64_t)k) && (!t || t[n] == (int64_t)bb_value);
    const unsigned m = __ballot_sync(P2C_FULL_MASK, in);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int off = base_s;
    int tot = 0;
#pragma unroll
    for (int w = 0; w < LIST_THREADS / 32; ++w) {
      if (w < warp) off += warp_tot[w];
      tot += warp_tot[w];
    }
    if (in && out) out[off + __popc(m & ((1u << lane) - 1u))] = n;
    __syncthreads();
    if (threadIdx.x == 0) base_s += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[(size_t)b * K + k] = base_s;
}

// torchgeometry 0.1.2 angle_axis_to_rotation_matrix (restated; see oracle/p2c_oracle.py) on aa = (a x z) * angle.
__device__ void rotation_to_z(const float* ax, float zero_tol, float R[9]) {
  R[0] = 1.f; R[1] = 0.f; R[2] = 0.f; R[3] = 0.f; R[4] = 1.f; R[5] = 0.f; R[6] = 0.f; R[7] = 0.f; R[8] = 1.f;
  const float angle = acosf(ax[2]);                       // <a, (0,0,1)>
  if (!(angle > zero_tol)) return;                        // NaN angle (|a_z| > 1) keeps the identity, like `if angle > tol`
  const float rx = ax[1] * angle, ry = -ax[0] * angle, rz = 0.f * angle;   // cross(a, z) * angle
  const float theta2 = rx * rx + ry * ry + rz * rz;
  if (theta2 > 1e-6f) {
    const float theta = sqrtf(theta2);
    const float wx = rx / (theta + 1e-6f), wy = ry / (theta + 1e-6f), wz = rz / (theta + 1e-6f);
    const float c = cosf(theta), s = sinf(theta), one_c = 1.f - c;
    R[0] = c + wx * wx * one_c;
    R[3] = wz * s + wx * wy * one_c;
    R[6] = -wy * s + wx * wz * one_c;
    R[1] = wx * wy * one_c - wz * s;
    R[4] = c + wy * wy * one_c;
    R[7] = wx * s + wy * wz * one_c;
    R[2] = wy * s + wx * wz * one_c;
    R[5] = -wx * s + wy * wz * one_c;
    R[8] = c + wz * wz * one_c;
  } else {
    R[0] = 1.f; R[1] = -rz; R[2] = ry; R[3] = rz; R[4] = 1.f; R[5] = -rx; R[6] = -ry; R[7] = rx; R[8] = 1.f;
  }
}

// (segment has > 1 member over the whole batch, this cloud has > 1) — the two `continue`s of the reference loop
__device__ void found_flags(const int32_t* counts, int B, int K, int b, int k, bool* seg_active, bool* found) {
  int64_t tot = 0;
  for (int j = 0; j < B; ++j) tot += counts[(size_t)j * K + k];
  *seg_active = tot > 1;
  *found = *seg_active && counts[(size_t)b * K + k] > 1;
}

__device__ __forceinline__ int64_t pick_member(const int32_t* list, const int64_t* rnd, int s, int cnt) {
  int64_t r = rnd ? rnd[s] : (int64_t)s;
  if (r < 0 || r >= cnt) return -1;
  return list ? (int64_t)list[r] : r;
}

__global__ void __launch_bounds__(256)
sketch_project_kernel(const float* __restrict__ P, const float* __restrict__ X, int B, int N, int K, int S,
                      const int32_t* __restrict__ lists, const int32_t* __restrict__ counts,
                      const int64_t* __restrict__ rand_idx, const float* __restrict__ axes,
                      const float* __restrict__ centers, float zero_tol, float* __restrict__ P_proj,
                      float* __restrict__ X_proj, float* __restrict__ scales, float* __restrict__ found_out,
                      int32_t* __restrict__ sel_out, float* __restrict__ R_out) {
  const int k = blockIdx.x, b = blockIdx.y;
  __shared__ float R[9];
  __shared__ float cproj[2];
  __shared__ bool flags[2];
  __shared__ float wmax[8];
  if (threadIdx.x == 0) {
    found_flags(counts, B, K, b, k, &flags[0], &flags[1]);
    const float* a = axes + ((size_t)b * K + k) * 3;
    const float ax[3] = {a[0], a[1], a[2]};
    float r[9];
    rotation_to_z(ax, zero_tol, r);
    for (int i = 0; i < 9; ++i) R[i] = r[i];
    const float* c = centers + ((size_t)b * K + k) * 3;
    cproj[0] = c[0] * r[0] + c[1] * r[3] + c[2] * r[6];
    cproj[1] = c[0] * r[1] + c[1] * r[4] + c[2] * r[7];
    if (R_out)
      for (int i = 0; i < 9; ++i) R_out[((size_t)k * B + b) * 9 + i] = r[i];
  }
  __syncthreads();
  const bool seg_active = flags[0], found = flags[1];
  const size_t ob = ((size_t)k * B + b) * S;
  const int cnt = counts[(size_t)b * K + k];
  const int32_t* list = lists ? lists + ((size_t)b * K + k) * N : nullptr;
  const int64_t* rnd = rand_idx ? rand_idx + ob : nullptr;
  const float* Pb = P + (size_t)b * N * 3;
  const float* Xb = X ? X + (size_t)b * N * 3 : nullptr;
  float best = 0.f;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    float px = 0.f, py = 0.f, nx = 0.f, ny = 0.f;
    int32_t chosen = -1;
    if (seg_active) {
      float p0 = 0.f, p1 = 0.f, p2 = 0.f, x0 = 0.f, x1 = 0.f, x2 = 0.f;
      if (found) {
        const int64_t n = pick_member(list, rnd, s, cnt);
        if (n >= 0 && n < N) {
          chosen = (int32_t)n;
          p0 = Pb[n * 3]; p1 = Pb[n * 3 + 1]; p2 = Pb[n * 3 + 2];
          if (Xb) { x0 = Xb[n * 3]; x1 = Xb[n * 3 + 1]; x2 = Xb[n * 3 + 2]; }
        }
      }
      px = (p0 * R[0] + p1 * R[3] + p2 * R[6]) - cproj[0];
      py = (p0 * R[1] + p1 * R[4] + p2 * R[7]) - cproj[1];
      nx = x0 * R[0] + x1 * R[3] + x2 * R[6];
      ny = x0 * R[1] + x1 * R[4] + x2 * R[7];
    }
    P_proj[(ob + s) * 2] = px;
    P_proj[(ob + s) * 2 + 1] = py;
    if (X_proj) {
      X_proj[(ob + s) * 2] = nx;
      X_proj[(ob + s) * 2 + 1] = ny;
    }
    if (sel_out) sel_out[ob + s] = chosen;
    best = fmaxf(best, sqrtf(px * px + py * py));
  }
  best = p2c_warp_max(best);
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = wmax[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, wmax[w]);
    scales[(size_t)k * B + b] = found ? m : 1.f;
    if (found_out) found_out[(size_t)b * K + k] = found ? 1.f : 0.f;
  }
}

__global__ void __launch_bounds__(256)
extents_kernel(const float* __restrict__ P, int B, int N, int K, int S, const int32_t* __restrict__ lists,
               const int32_t* __restrict__ counts, const int64_t* __restrict__ rand_idx,
               const float* __restrict__ axes, const float* __restrict__ centers, float* __restrict__ extents,
               float* __restrict__ found_out) {
  const int k = blockIdx.x, b = blockIdx.y;
  __shared__ bool flags[2];
  __shared__ float wmin[8], wmax[8];
  if (threadIdx.x == 0) found_flags(counts, B, K, b, k, &flags[0], &flags[1]);
  __syncthreads();
  const bool seg_active = flags[0], found = flags[1];
  const float* a = axes + ((size_t)b * K + k) * 3;
  const float* c = centers + ((size_t)b * K + k) * 3;
  const float a0 = a[0], a1 = a[1], a2 = a[2], c0 = c[0], c1 = c[1], c2 = c[2];
  const size_t ob = ((size_t)k * B + b) * S;
  const int cnt = counts[(size_t)b * K + k];
  const int32_t* list = lists ? lists + ((size_t)b * K + k) * N : nullptr;
  const int64_t* rnd = rand_idx ? rand_idx + ob : nullptr;
  const float* Pb = P + (size_t)b * N * 3;
  float lo = INFINITY, hi = -INFINITY;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    float p0 = 0.f, p1 = 0.f, p2 = 0.f;
    if (found) {
      const int64_t n = pick_member(list, rnd, s, cnt);
      if (n >= 0 && n < N) { p0 = Pb[n * 3]; p1 = Pb[n * 3 + 1]; p2 = Pb[n * 3 + 2]; }
    }
    const float d = (p0 - c0) * a0 + (p1 - c1) * a1 + (p2 - c2) * a2;
    lo = fminf(lo, d);
    hi = fmaxf(hi, d);
  }
  hi = p2c_warp_max(hi);
  lo = -p2c_warp_max(-lo);
  if ((threadIdx.x & 31) == 0) { wmin[threadIdx.x >> 5] = lo; wmax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = fminf(lo, wmin[w]); hi = fmaxf(hi, wmax[w]); }
    extents[((size_t)k * B + b) * 2] = seg_active ? lo : 0.f;
    extents[((size_t)k * B + b) * 2 + 1] = seg_active ? hi : 0.f;
    if (found_out) found_out[(size_t)b * K + k] = found ? 1.f : 0.f;
  }
}

// backward of the projected normals w.r.t. X: X_proj[k,b,s,c] = sum_r X[b, sel, r] * R[r][c]  =>
// dX[b, sel, r] += R[r][0] * g0 + R[r][1] * g1   (the same point can be sampled many times: atomics)
__global__ void __launch_bounds__(256)
sketch_project_bwd_kernel(const float* __restrict__ dXproj, const int32_t* __restrict__ sel,
                          const float* __restrict__ R, int B, int N, int S, int64_t total, float* __restrict__ dX) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int n = sel[e];
  if (n < 0) return;
  const int64_t kb = e / S;                 // k*B + b
  const int b = (int)(kb % B);
  const float* r = R + kb * 9;
  const float g0 = dXproj[e * 2], g1 = dXproj[e * 2 + 1];
  float* o = dX + ((size_t)b * N + n) * 3;
  atomicAdd(o + 0, r[0] * g0 + r[1] * g1);
  atomicAdd(o + 1, r[3] * g0 + r[4] * g1);
  atomicAdd(o + 2, r[6] * g0 + r[7] * g1);
}

// hard_W_encoding: one-hot of the first arg-max over K, nulled columns zeroed; also the arg-max label itself
__global__ void __launch_bounds__(256)
hard_w_kernel(const float* __restrict__ W, int64_t ldw, int64_t sw, int64_t rows, int N, int K,
              const float* __restrict__ colsum, float null_below, float* __restrict__ hard,
              int64_t* __restrict__ label) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* w = W + r * ldw;
  int best = 0;
  float bv = w[0];
  for (int k = 1; k < K; ++k) {
    const float v = w[k * sw];
    if (v > bv || (v != v && bv == bv)) { bv = v; best = k; }   // first max; NaN counts as max like torch.argmax
  }
  if (label) label[r] = best;
  if (hard) {
    const int64_t b = r / N;
    float keep = 1.f;
    if (colsum) keep = colsum[b * K + best] < null_below ? 0.f : 1.f;
    float* h = hard + r * K;
    for (int k = 0; k < K; ++k) h[k] = (k == best) ? keep : 0.f;
  }
}

// compute_normal_difference / compute_normal_loss(angle_diff): acos_safe(|<x, g>|) per point, optional per-cloud sum
__global__ void __launch_bounds__(256)
normal_angle_kernel(const float* __restrict__ X, const float* __restrict__ G, int N, float scale,
                    float* __restrict__ per_point, float* __restrict__ per_cloud_sum) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  float v = 0.f;
  if (n < N) {
    const float* x = X + ((size_t)b * N + n) * 3;
    const float* g = G + ((size_t)b * N + n) * 3;
    float d = fabsf(x[0] * g[0] + x[1] * g[1] + x[2] * g[2]);
    d = fminf(fmaxf(d, -1.0f + 1e-6f), 1.0f - 1e-6f);
    v = acosf(d) * scale;
    if (per_point) per_point[(size_t)b * N + n] = v;
  }
  if (per_cloud_sum) {
    __shared__ float ws[8];
    v = p2c_warp_sum(v);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += ws[w];
      atomicAdd(per_cloud_sum + b, t);
    }
  }
}

}  // namespace

extern "C" int p2c_segment_lists(const int64_t* seg_label, const int64_t* bb, int bb_value, int B, int N, int K,
                                 int32_t* counts, int32_t* lists, void* stream) {
  if (!seg_label || !counts || B <= 0 || N <= 0 || K <= 0 || B > 65535) return P2C_EINVAL;
  segment_lists_kernel<<<dim3(K, B), LIST_THREADS, 0, (cudaStream_t)stream>>>(seg_label, bb, bb_value, N, K, counts,
                                                                              lists);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_sketch_project(const float* P, const float* X, int B, int N, int K, int S, const int32_t* lists,
                                  const int32_t* counts, const int64_t* rand_idx, const float* axes,
                                  const float* centers, float zero_tol, float* P_proj, float* X_proj, float* scales,
                                  float* found, int32_t* sel_out, float* R_out, void* stream) {
  if (!P || !counts || !axes || !centers || !P_proj || !scales || B <= 0 || N <= 0 || K <= 0 || S <= 0 || B > 65535)
    return P2C_EINVAL;
  if ((X == nullptr) != (X_proj == nullptr)) return P2C_EINVAL;
  sketch_project_kernel<<<dim3(K, B), 256, 0, (cudaStream_t)stream>>>(P, X, B, N, K, S, lists, counts, rand_idx, axes,
                                                                      centers, zero_tol, P_proj, X_proj, scales, found,
                                                                      sel_out, R_out);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_sketch_project_bwd(const float* dX_proj, const int32_t* sel, const float* R, int B, int N, int K, int S,
                                      float* dX, void* stream) {
  if (!dX_proj || !sel || !R || !dX || B <= 0 || N <= 0 || K <= 0 || S <= 0) return P2C_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  P2C_CUDA_TRY(cudaMemsetAsync(dX, 0, sizeof(float) * (size_t)B * N * 3, st));
  const int64_t total = (int64_t)K * B * S;
  sketch_project_bwd_kernel<<<p2c_ceil_div(total, 256), 256, 0, st>>>(dX_proj, sel, R, B, N, S, total, dX);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_extrusion_extents(const float* P, int B, int N, int K, int S, const int32_t* lists,
                                     const int32_t* counts, const int64_t* rand_idx, const float* axes,
                                     const float* centers, float* extents, float* found, void* stream) {
  if (!P || !counts || !axes || !centers || !extents || B <= 0 || N <= 0 || K <= 0 || S <= 0 || B > 65535)
    return P2C_EINVAL;
  extents_kernel<<<dim3(K, B), 256, 0, (cudaStream_t)stream>>>(P, B, N, K, S, lists, counts, rand_idx, axes, centers,
                                                               extents, found);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_hard_w_encoding(const float* W, int64_t ldw, int64_t sw, int B, int N, int K, const float* colsum,
                                   float null_below, float* hard, int64_t* label, void* stream) {
  if (!W || (!hard && !label) || B <= 0 || N <= 0 || K <= 0) return P2C_EINVAL;
  const int64_t rows = (int64_t)B * N;
  hard_w_kernel<<<p2c_ceil_div(rows, 256), 256, 0, (cudaStream_t)stream>>>(W, ldw, sw, rows, N, K, colsum, null_below,
                                                                          hard, label);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_normal_angle(const float* X, const float* G, int B, int N, float scale, float* per_point,
                                float* per_cloud_sum, void* stream) {
  if (!X || !G || (!per_point && !per_cloud_sum) || B <= 0 || N <= 0 || B > 65535) return P2C_EINVAL;
  if (per_cloud_sum) P2C_CUDA_TRY(cudaMemsetAsync(per_cloud_sum, 0, sizeof(float) * B, (cudaStream_t)stream));
  normal_angle_kernel<<<dim3(p2c_ceil_div(N, 256), B), 256, 0, (cudaStream_t)stream>>>(X, G, N, scale, per_point,
                                                                                      per_cloud_sum);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
