// Loss-side kernels: one pass over the points produces every per-cloud sufficient statistic of the
// Point2Cyl loss block (train_Point2Cyl_without_sketch.py:246-353), a second light pass produces the
// base/barrel loss once the Hungarian match is known, and a finalisation kernel turns the
// statistics into relaxed IoU, centres, the per-segment 3x3 eigen-solve and the masked means.
//
// Why this is enough (DESIGN.md has the algebra):
//   * relaxed IoU (losses.py:95-101) and the Hungarian cost (losses.py:39-42) only need
//     D[g,k] = sum_n [inst=g] W[n,k], cnt[g], colsum[k];
//   * centres (data_utils.py:253-266) need C[k] = sum_n W[n,k] p_n;
//   * BtB - CtC (data_utils.py:155-163) is sum_n (wbar^2 - wbase^2) x x^T: 6+6 moments per
//     predicted column; re-ordering columns by the match is a gather of these small results, so
//     the reference's (B,N,N) diag_embed matrices never exist;
//   * the sort in the inline bb loss (train_...:292) permutes a sum over slots and cancels.
//
// Thread layout of the point passes: a group of KP (= K rounded up to a power of two) adjacent
// lanes owns one point, lane k of the group owns predicted column k (barrel logit 2k, base 2k+1).
//
// stats layout per cloud (float32), K columns, stride K*K + 19*K + 2:
//   D[K][K] | cnt[K] | cnt_barrel[K] | cnt_base[K] | colsum[K] | C[K][3] | Mbar[K][6] | Mbase[K][6]
//   | normal_sum | max_label
#include "segfit_common.cuh"

namespace {


template <int KP>
__global__ void __launch_bounds__(SEG_THREADS)
segfit_partial_kernel(const float* __restrict__ X_raw, int64_t ldx, const float* __restrict__ W_raw,
                      int64_t ldw, const float* __restrict__ pcs, const float* __restrict__ gtn,
                      const int64_t* __restrict__ inst, const int64_t* __restrict__ bb, int N, int K,
                      float* __restrict__ partial, int nchunks) {
  constexpr int NACC = KP + 21;
  constexpr int PPS = SEG_THREADS / KP;  // points per step
  __shared__ float s_red[SEG_THREADS / 32][KP][NACC];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tid = threadIdx.x;
  const int k = tid % KP, pl = tid / KP;
  const int lane = tid & 31, warp = tid >> 5;

  float accD[KP];
#pragma unroll
  for (int g = 0; g < KP; ++g) accD[g] = 0.f;
  float colsum = 0.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
  float mb[6] = {0, 0, 0, 0, 0, 0}, mc[6] = {0, 0, 0, 0, 0, 0};
  float cnt = 0.f, cbar = 0.f, cbase = 0.f, nsum = 0.f, maxlab = -1.f;

  const int n_end = min(N, (chunk + 1) * SEG_CHUNK);
  for (int n0 = chunk * SEG_CHUNK; n0 < n_end; n0 += PPS) {
    // whole warps stay converged (the shuffles below are warp-wide); out-of-range points add 0
    const int n = n0 + pl;
    const bool ok = n < n_end;
    const size_t row = (size_t)b * N + (ok ? n : (n_end - 1));
    float rbar, rbase, wb, wc;
    point_softmax<KP>(W_raw + row * ldw, k, K, rbar, rbase, wb, wc);
    const float* xr = X_raw + row * ldx;
    float x = __ldg(xr), y = __ldg(xr + 1), z = __ldg(xr + 2);
    const float inv = 1.0f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);  // F.normalize eps
    x *= inv; y *= inv; z *= inv;
    const int g = (int)inst[row];
    const int t = (int)bb[row];
    const float valid = ok ? 1.f : 0.f;
    wb *= valid; wc *= valid;
    const float w = wb + wc;
#pragma unroll
    for (int gg = 0; gg < KP; ++gg) accD[gg] += (g == gg) ? w : 0.f;
    colsum += w;
    const float* pr = pcs + row * 3;
    C0 = fmaf(w, __ldg(pr), C0); C1 = fmaf(w, __ldg(pr + 1), C1); C2 = fmaf(w, __ldg(pr + 2), C2);
    const float b2 = wb * wb, c2 = wc * wc;
    const float xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
    mb[0] = fmaf(b2, xx, mb[0]); mb[1] = fmaf(b2, xy, mb[1]); mb[2] = fmaf(b2, xz, mb[2]);
    mb[3] = fmaf(b2, yy, mb[3]); mb[4] = fmaf(b2, yz, mb[4]); mb[5] = fmaf(b2, zz, mb[5]);
    mc[0] = fmaf(c2, xx, mc[0]); mc[1] = fmaf(c2, xy, mc[1]); mc[2] = fmaf(c2, xz, mc[2]);
    mc[3] = fmaf(c2, yy, mc[3]); mc[4] = fmaf(c2, yz, mc[4]); mc[5] = fmaf(c2, zz, mc[5]);
    if (ok) {
      if (g == k) { cnt += 1.f; cbar += (t == 0) ? 1.f : 0.f; cbase += (t == 1) ? 1.f : 0.f; }
      if (k == 0) {
        const float* gr = gtn + row * 3;
        nsum += 1.0f - fabsf(x * __ldg(gr) + y * __ldg(gr + 1) + z * __ldg(gr + 2));
        maxlab = fmaxf(maxlab, (float)g);
      }
    }
  }

  // reduce over the point dimension: lanes with equal k inside the warp, then across warps
  float acc[NACC];
#pragma unroll
  for (int g = 0; g < KP; ++g) acc[g] = accD[g];
  acc[KP] = colsum; acc[KP + 1] = C0; acc[KP + 2] = C1; acc[KP + 3] = C2;
#pragma unroll
  for (int i = 0; i < 6; ++i) { acc[KP + 4 + i] = mb[i]; acc[KP + 10 + i] = mc[i]; }
  acc[KP + 16] = cnt; acc[KP + 17] = cbar; acc[KP + 18] = cbase; acc[KP + 19] = nsum;
#pragma unroll
  for (int i = 0; i < KP + 20; ++i) acc[i] = column_sum<KP>(acc[i]);
#pragma unroll
  for (int o = 16; o >= KP; o >>= 1) maxlab = fmaxf(maxlab, __shfl_xor_sync(P2C_FULL_MASK, maxlab, o));
  acc[KP + 20] = maxlab;
  if (lane < KP) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) s_red[warp][lane][i] = acc[i];
  }
  __syncthreads();
  if (tid < KP && tid < K) {
    float tot[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) tot[i] = 0.f;
    tot[KP + 20] = -1.f;
    for (int w = 0; w < SEG_THREADS / 32; ++w) {
#pragma unroll
      for (int i = 0; i < KP + 20; ++i) tot[i] += s_red[w][tid][i];
      tot[KP + 20] = fmaxf(tot[KP + 20], s_red[w][tid][KP + 20]);
    }
    float* o = partial + ((size_t)b * nchunks + chunk) * seg_stride(K);
    const int kk = tid;
#pragma unroll
    for (int g = 0; g < KP; ++g)
      if (g < K) o[g * K + kk] = tot[g];
    o[off_colsum(K) + kk] = tot[KP];
    o[off_C(K) + kk * 3 + 0] = tot[KP + 1];
    o[off_C(K) + kk * 3 + 1] = tot[KP + 2];
    o[off_C(K) + kk * 3 + 2] = tot[KP + 3];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      o[off_Mbar(K) + kk * 6 + i] = tot[KP + 4 + i];
      o[off_Mbase(K) + kk * 6 + i] = tot[KP + 10 + i];
    }
    o[off_cnt(K) + kk] = tot[KP + 16];
    o[off_cbar(K) + kk] = tot[KP + 17];
    o[off_cbase(K) + kk] = tot[KP + 18];
    if (kk == 0) {
      o[off_normal(K)] = tot[KP + 19];
      o[off_maxlab(K)] = tot[KP + 20];
    }
  }
}

// Function-level variant (drop-in losses.py / data_utils.py API): the caller already holds the soft assignments.
// wb/wc: (B,N,K) with row stride ld* and element stride s* (slices like W_2K[:, :, ::2] are accepted); wc, X,
// pcs, gt normals and bb may be NULL - the corresponding statistics are then zero.  X is used as given
// (normalize_x = 0) or L2-normalised like F.normalize (normalize_x = 1).
template <int KP>
__global__ void __launch_bounds__(SEG_THREADS)
segfit_partial_w_kernel(const float* __restrict__ X, int64_t ldx, int normalize_x, const float* __restrict__ wbp,
                        int64_t ldb, int64_t sb, const float* __restrict__ wcp, int64_t ldc, int64_t sc,
                        const float* __restrict__ pcs, const float* __restrict__ gtn,
                        const int64_t* __restrict__ inst, const int64_t* __restrict__ bb, int N, int K,
                        float* __restrict__ partial, int nchunks) {
  constexpr int NACC = KP + 21;
  constexpr int PPS = SEG_THREADS / KP;
  __shared__ float s_red[SEG_THREADS / 32][KP][NACC];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tid = threadIdx.x;
  const int k = tid % KP, pl = tid / KP;
  const int lane = tid & 31, warp = tid >> 5;
  float accD[KP];
#pragma unroll
  for (int g = 0; g < KP; ++g) accD[g] = 0.f;
  float colsum = 0.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
  float mb[6] = {0, 0, 0, 0, 0, 0}, mc[6] = {0, 0, 0, 0, 0, 0};
  float cnt = 0.f, cbar = 0.f, cbase = 0.f, nsum = 0.f, maxlab = -1.f;
  const int n_end = min(N, (chunk + 1) * SEG_CHUNK);
  for (int n0 = chunk * SEG_CHUNK; n0 < n_end; n0 += PPS) {
    const int n = n0 + pl;
    const bool ok = n < n_end;
    const size_t row = (size_t)b * N + (ok ? n : (n_end - 1));
    float wb = 0.f, wc = 0.f;
    if (ok && k < K) {
      wb = __ldg(wbp + row * ldb + (size_t)k * sb);
      if (wcp) wc = __ldg(wcp + row * ldc + (size_t)k * sc);
    }
    float x = 0.f, y = 0.f, z = 0.f;
    if (X) {
      const float* xr = X + row * ldx;
      x = __ldg(xr); y = __ldg(xr + 1); z = __ldg(xr + 2);
      if (normalize_x) {
        const float inv = 1.0f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
        x *= inv; y *= inv; z *= inv;
      }
    }
    const int g = inst ? (int)inst[row] : -1;
    const int t = bb ? (int)bb[row] : -1;
    const float w = wb + wc;
#pragma unroll
    for (int gg = 0; gg < KP; ++gg) accD[gg] += (g == gg) ? w : 0.f;
    colsum += w;
    if (pcs) {
      const float* pr = pcs + row * 3;
      C0 = fmaf(w, __ldg(pr), C0); C1 = fmaf(w, __ldg(pr + 1), C1); C2 = fmaf(w, __ldg(pr + 2), C2);
    }
    const float b2 = wb * wb, c2 = wc * wc;
    const float xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
    mb[0] = fmaf(b2, xx, mb[0]); mb[1] = fmaf(b2, xy, mb[1]); mb[2] = fmaf(b2, xz, mb[2]);
    mb[3] = fmaf(b2, yy, mb[3]); mb[4] = fmaf(b2, yz, mb[4]); mb[5] = fmaf(b2, zz, mb[5]);
    mc[0] = fmaf(c2, xx, mc[0]); mc[1] = fmaf(c2, xy, mc[1]); mc[2] = fmaf(c2, xz, mc[2]);
    mc[3] = fmaf(c2, yy, mc[3]); mc[4] = fmaf(c2, yz, mc[4]); mc[5] = fmaf(c2, zz, mc[5]);
    if (ok) {
      if (g == k) { cnt += 1.f; cbar += (t == 0) ? 1.f : 0.f; cbase += (t == 1) ? 1.f : 0.f; }
      if (k == 0) {
        if (gtn && X) {
          const float* gr = gtn + row * 3;
          nsum += 1.0f - fabsf(x * __ldg(gr) + y * __ldg(gr + 1) + z * __ldg(gr + 2));
        }
        maxlab = fmaxf(maxlab, (float)g);
      }
    }
  }
  float acc[NACC];
#pragma unroll
  for (int g = 0; g < KP; ++g) acc[g] = accD[g];
  acc[KP] = colsum; acc[KP + 1] = C0; acc[KP + 2] = C1; acc[KP + 3] = C2;
#pragma unroll
  for (int i = 0; i < 6; ++i) { acc[KP + 4 + i] = mb[i]; acc[KP + 10 + i] = mc[i]; }
  acc[KP + 16] = cnt; acc[KP + 17] = cbar; acc[KP + 18] = cbase; acc[KP + 19] = nsum;
#pragma unroll
  for (int i = 0; i < KP + 20; ++i) acc[i] = column_sum<KP>(acc[i]);
#pragma unroll
  for (int o = 16; o >= KP; o >>= 1) maxlab = fmaxf(maxlab, __shfl_xor_sync(P2C_FULL_MASK, maxlab, o));
  acc[KP + 20] = maxlab;
  if (lane < KP) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) s_red[warp][lane][i] = acc[i];
  }
  __syncthreads();
  if (tid < KP && tid < K) {
    float tot[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) tot[i] = 0.f;
    tot[KP + 20] = -1.f;
    for (int w = 0; w < SEG_THREADS / 32; ++w) {
#pragma unroll
      for (int i = 0; i < KP + 20; ++i) tot[i] += s_red[w][tid][i];
      tot[KP + 20] = fmaxf(tot[KP + 20], s_red[w][tid][KP + 20]);
    }
    float* o = partial + ((size_t)b * nchunks + chunk) * seg_stride(K);
    const int kk = tid;
#pragma unroll
    for (int g = 0; g < KP; ++g)
      if (g < K) o[g * K + kk] = tot[g];
    o[off_colsum(K) + kk] = tot[KP];
    o[off_C(K) + kk * 3 + 0] = tot[KP + 1];
    o[off_C(K) + kk * 3 + 1] = tot[KP + 2];
    o[off_C(K) + kk * 3 + 2] = tot[KP + 3];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      o[off_Mbar(K) + kk * 6 + i] = tot[KP + 4 + i];
      o[off_Mbase(K) + kk * 6 + i] = tot[KP + 10 + i];
    }
    o[off_cnt(K) + kk] = tot[KP + 16];
    o[off_cbar(K) + kk] = tot[KP + 17];
    o[off_cbase(K) + kk] = tot[KP + 18];
    if (kk == 0) {
      o[off_normal(K)] = tot[KP + 19];
      o[off_maxlab(K)] = tot[KP + 20];
    }
  }
}

// fixed-order sum of the chunk partials (deterministic); max for the label slot
__global__ void segfit_reduce_kernel(const float* __restrict__ partial, int nchunks, int K,
                                     float* __restrict__ stats) {
  const int b = blockIdx.x;
  const int stride = seg_stride(K);
  for (int i = threadIdx.x; i < stride; i += blockDim.x) {
    const float* p = partial + (size_t)b * nchunks * stride + i;
    if (i == off_maxlab(K)) {
      float m = -1.f;
      for (int c = 0; c < nchunks; ++c) m = fmaxf(m, p[(size_t)c * stride]);
      stats[(size_t)b * stride + i] = m;
    } else {
      // four independent chains, eight loads in flight: with 32 chunks per cloud (256-point chunks) a single
      // dependent load-add chain made this 23 us of pure latency
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int c = 0;
      for (; c + 8 <= nchunks; c += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(p + (size_t)(c + u) * stride);
        s0 += (double)v[0] + (double)v[4]; s1 += (double)v[1] + (double)v[5];
        s2 += (double)v[2] + (double)v[6]; s3 += (double)v[3] + (double)v[7];
      }
      for (; c < nchunks; ++c) s0 += (double)__ldg(p + (size_t)c * stride);
      stats[(size_t)b * stride + i] = (float)((s0 + s1) + (s2 + s3));
    }
  }
}

__global__ void segfit_cost_kernel(const float* __restrict__ stats, int B, int K,
                                   float* __restrict__ cost, int32_t* __restrict__ n_gt) {
  const int b = blockIdx.x;
  const float* s = stats + (size_t)b * seg_stride(K);
  for (int e = threadIdx.x; e < K * K; e += blockDim.x) {
    const int g = e / K, k = e % K;
    const float d = s[g * K + k];
    const float den = s[off_cnt(K) + g] + s[off_colsum(K) + k] - d;
    cost[(size_t)b * K * K + e] = d / fmaxf(den, 1e-10f);
  }
  if (threadIdx.x == 0) n_gt[b] = (int)s[off_maxlab(K)] + 1;
}

template <int KP>
__global__ void __launch_bounds__(SEG_THREADS)
bb_partial_kernel(const float* __restrict__ W_raw, int64_t ldw, const int64_t* __restrict__ bb,
                  const int64_t* __restrict__ match, const int32_t* __restrict__ n_gt, int N, int K,
                  float* __restrict__ partial, int nchunks) {
  constexpr int PPS = SEG_THREADS / KP;
  __shared__ float s_red[SEG_THREADS / 32];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tid = threadIdx.x;
  const int k = tid % KP, pl = tid / KP;
  const int lane = tid & 31, warp = tid >> 5;
  const int my_col = k < K ? (int)match[(size_t)b * K + k] : 0;  // slot k takes predicted column my_col
  const bool slot_live = k < K && k < n_gt[b];
  float acc = 0.f;
  const int n_end = min(N, (chunk + 1) * SEG_CHUNK);
  for (int n0 = chunk * SEG_CHUNK; n0 < n_end; n0 += PPS) {
    const int n = n0 + pl;
    const bool ok = n < n_end;
    const size_t row = (size_t)b * N + (ok ? n : (n_end - 1));
    float rbar, rbase, wb, wc;
    point_softmax<KP>(W_raw + row * ldw, k, K, rbar, rbase, wb, wc);
    const float w = wb + wc;
    // W re-ordered by the match, unmatched slots zeroed (train_...:287-289), softmax over slots
    const float wm = __shfl_sync(P2C_FULL_MASK, w, (lane & ~(KP - 1)) + my_col);
    const float zin = slot_live ? wm : 0.f;
    const float NEG = -__int_as_float(0x7f800000);
    const float zmax = group_max<KP>(k < K ? zin : NEG);
    const float ez = k < K ? expf(zin - zmax) : 0.f;
    const float z = ez / group_sum<KP>(ez);
    // 2-way cross entropy on the raw logits of column k (not re-ordered: reference quirk a14)
    float ce = 0.f;
    if (k < K) {
      const float mx = fmaxf(rbar, rbase);
      const float lse = mx + logf(expf(rbar - mx) + expf(rbase - mx));
      ce = lse - ((int)bb[row] == 0 ? rbar : rbase);
    }
    const float tot = group_sum<KP>(z * ce);
    if (ok && k == 0) acc += tot;
  }
  acc = p2c_warp_sum(acc);
  if (lane == 0) s_red[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < SEG_THREADS / 32; ++w) s += s_red[w];
    partial[(size_t)b * nchunks + chunk] = s;
  }
}

__global__ void bb_reduce_kernel(const float* __restrict__ partial, int nchunks, int B,
                                 float* __restrict__ bb_sum) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double s = 0.0;
  for (int c = 0; c < nchunks; ++c) s += (double)partial[(size_t)b * nchunks + c];
  bb_sum[b] = (float)s;
}


__global__ void eig3x3_kernel(const float* __restrict__ M, int n, float* __restrict__ vec,
                              float* __restrict__ eval) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* m = M + (size_t)i * 9;
  // torch.symeig(upper=True): only the upper triangle is read (data_utils.py:170)
  const double a[6] = {m[0], m[1], m[2], m[4], m[5], m[8]};
  double e[3], v[3][3];
  jacobi3(a, e, v);
  vec[i * 3 + 0] = (float)v[0][0]; vec[i * 3 + 1] = (float)v[0][1]; vec[i * 3 + 2] = (float)v[0][2];
  if (eval) { eval[i * 3 + 0] = (float)e[0]; eval[i * 3 + 1] = (float)e[1]; eval[i * 3 + 2] = (float)e[2]; }
}

// one warp per (cloud, gt slot); lane 0 runs the Jacobi solve.  per_seg: (B,K,3) {miou, axis, center}
__global__ void __launch_bounds__(128)
loss_segments_kernel(const float* __restrict__ stats, const int64_t* __restrict__ match,
                     const int32_t* __restrict__ n_gt, const float* __restrict__ gt_axes,
                     const float* __restrict__ gt_centers, int B, int N, int K, int norm_eig,
                     float* __restrict__ E_AX, float* __restrict__ centers,
                     float* __restrict__ per_seg) {
  const int seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (seg >= B * K) return;
  const int b = seg / K, j = seg % K;
  const float* s = stats + (size_t)b * seg_stride(K);
  const int m = (int)match[seg];
  // lanes 0..5 fetch one moment each; the solve itself is serial
  float mbar = 0.f, mbase = 0.f;
  if (lane < 6) { mbar = s[off_Mbar(K) + m * 6 + lane]; mbase = s[off_Mbase(K) + m * 6 + lane]; }
  double sb = 1.0, sc = 1.0;
  if (norm_eig) {
    const double nb = sqrt((double)s[off_cbar(K) + j]) + 1.0, nc = sqrt((double)s[off_cbase(K) + j]) + 1.0;
    sb = 1.0 / (nb * nb); sc = 1.0 / (nc * nc);
  }
  const double mval = (double)mbar * sb - (double)mbase * sc;
  double a[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) a[i] = __shfl_sync(P2C_FULL_MASK, mval, i);
  if (lane == 0) {
    double e[3], v[3][3];
    jacobi3(a, e, v);
    const float ax = (float)v[0][0], ay = (float)v[0][1], az = (float)v[0][2];
    E_AX[seg * 3 + 0] = ax; E_AX[seg * 3 + 1] = ay; E_AX[seg * 3 + 2] = az;
    const float* ga = gt_axes + (size_t)seg * 3;
    const float l_ax = 1.0f - fabsf(ax * ga[0] + ay * ga[1] + az * ga[2]);
    const float invN = 1.0f / (float)N;
    const float c0 = s[off_C(K) + m * 3 + 0] * invN, c1 = s[off_C(K) + m * 3 + 1] * invN,
                c2 = s[off_C(K) + m * 3 + 2] * invN;
    centers[seg * 3 + 0] = c0; centers[seg * 3 + 1] = c1; centers[seg * 3 + 2] = c2;
    const float* gc = gt_centers + (size_t)seg * 3;
    const float d0 = c0 - gc[0], d1 = c1 - gc[1], d2 = c2 - gc[2];
    const float l_c = d0 * d0 + d1 * d1 + d2 * d2;
    const float dj = s[j * K + m];
    const float iou = dj / (s[off_cnt(K) + j] + s[off_colsum(K) + m] - dj + 1e-10f);
    per_seg[seg * 3 + 0] = 1.0f - iou;
    per_seg[seg * 3 + 1] = l_ax;
    per_seg[seg * 3 + 2] = l_c;
  }
}

struct LossWeights { float w[5]; };

// single CTA: masked means per cloud (losses.py:83-88) then the batch means (losses.py:335-340)
__global__ void loss_reduce_kernel(const float* __restrict__ stats, const float* __restrict__ bb_sum,
                                   const int32_t* __restrict__ n_gt, const float* __restrict__ per_seg,
                                   int B, int N, int K, LossWeights lw, float* __restrict__ per_cloud,
                                   float* __restrict__ out) {
  __shared__ double s_tot[5];
  if (threadIdx.x < 5) s_tot[threadIdx.x] = 0.0;
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const int ng = min(n_gt[b], K);
    float miou = 0.f, ax = 0.f, ce = 0.f;
    for (int j = 0; j < ng; ++j) {
      miou += per_seg[(b * K + j) * 3 + 0];
      ax += per_seg[(b * K + j) * 3 + 1];
      ce += per_seg[(b * K + j) * 3 + 2];
    }
    const float den = ng > 0 ? (float)ng : 1.f;
    float* pc = per_cloud + (size_t)b * 5;
    pc[0] = miou / den;
    pc[1] = stats[(size_t)b * seg_stride(K) + off_normal(K)] / (float)N;
    pc[2] = bb_sum ? bb_sum[b] / (float)N : 0.f;
    pc[3] = ax / den;
    pc[4] = ce / den;
  }
  __syncthreads();
  if (threadIdx.x < 5) {  // fixed-order sums: deterministic
    double s = 0.0;
    for (int b = 0; b < B; ++b) s += (double)per_cloud[(size_t)b * 5 + threadIdx.x];
    s_tot[threadIdx.x] = s / (double)B;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // out: {total, normal, miou, bb, axis, center}; weights: {seg, normal, bb, extrusion, centre}
    const double miou = s_tot[0], nrm = s_tot[1], bbl = s_tot[2], ax = s_tot[3], ce = s_tot[4];
    out[0] = (float)(lw.w[0] * miou + lw.w[1] * nrm + lw.w[2] * bbl + lw.w[3] * ax + lw.w[4] * ce);
    out[1] = (float)nrm; out[2] = (float)miou; out[3] = (float)bbl; out[4] = (float)ax; out[5] = (float)ce;
  }
}

}  // namespace

extern "C" int p2c_segfit_stats_stride(int K) { return seg_stride(K); }

extern "C" int p2c_segfit_stats(const float* X_raw, int64_t ldx, const float* W_raw, int64_t ldw,
                                const float* pcs, const float* gt_normals, const int64_t* inst,
                                const int64_t* bb, int B, int N, int K, float* partial,
                                int64_t partial_elems, float* stats, void* stream) {
  if (!X_raw || !W_raw || !pcs || !gt_normals || !inst || !bb || !partial || !stats) return P2C_EINVAL;
  if (B <= 0 || N <= 0 || K <= 0 || ldx < 3 || ldw < 2 * K) return P2C_EINVAL;
  const int KP = kp_of(K);
  if (!KP) return P2C_EUNSUPPORTED;
  const int nchunks = p2c_ceil_div(N, SEG_CHUNK);
  if (partial_elems < (int64_t)B * nchunks * seg_stride(K)) return P2C_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(nchunks, B);
#define P2C_SEG_LAUNCH(KPV)                                                                     \
  segfit_partial_kernel<KPV><<<grid, SEG_THREADS, 0, st>>>(X_raw, ldx, W_raw, ldw, pcs, gt_normals, \
                                                           inst, bb, N, K, partial, nchunks)
  if (KP == 2) P2C_SEG_LAUNCH(2); else if (KP == 4) P2C_SEG_LAUNCH(4);
  else if (KP == 8) P2C_SEG_LAUNCH(8); else P2C_SEG_LAUNCH(16);
#undef P2C_SEG_LAUNCH
  P2C_RETURN_IF_CUDA_ERROR();
  segfit_reduce_kernel<<<B, 256, 0, st>>>(partial, nchunks, K, stats);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_segfit_stats_w(const float* X, int64_t ldx, int normalize_x, const float* wb, int64_t ldb,
                                  int64_t sb, const float* wc, int64_t ldc, int64_t sc, const float* pcs,
                                  const float* gt_normals, const int64_t* inst, const int64_t* bb, int B, int N,
                                  int K, float* partial, int64_t partial_elems, float* stats, void* stream) {
  if (!wb || !partial || !stats || B <= 0 || N <= 0 || K <= 0) return P2C_EINVAL;
  const int KP = kp_of(K);
  if (!KP) return P2C_EUNSUPPORTED;
  const int nchunks = p2c_ceil_div(N, SEG_CHUNK);
  if (partial_elems < (int64_t)B * nchunks * seg_stride(K)) return P2C_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(nchunks, B);
#define P2C_SEGW_LAUNCH(KPV)                                                                              \
  segfit_partial_w_kernel<KPV><<<grid, SEG_THREADS, 0, st>>>(X, ldx, normalize_x, wb, ldb, sb, wc, ldc, sc, pcs, \
                                                             gt_normals, inst, bb, N, K, partial, nchunks)
  if (KP == 2) P2C_SEGW_LAUNCH(2); else if (KP == 4) P2C_SEGW_LAUNCH(4);
  else if (KP == 8) P2C_SEGW_LAUNCH(8); else P2C_SEGW_LAUNCH(16);
#undef P2C_SEGW_LAUNCH
  P2C_RETURN_IF_CUDA_ERROR();
  segfit_reduce_kernel<<<B, 256, 0, st>>>(partial, nchunks, K, stats);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_segfit_cost(const float* stats, int B, int K, float* cost, int32_t* n_gt, void* stream) {
  if (!stats || !cost || !n_gt || B <= 0 || K <= 0) return P2C_EINVAL;
  segfit_cost_kernel<<<B, 64, 0, (cudaStream_t)stream>>>(stats, B, K, cost, n_gt);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_bb_loss(const float* W_raw, int64_t ldw, const int64_t* bb, const int64_t* match,
                           const int32_t* n_gt, int B, int N, int K, float* partial,
                           int64_t partial_elems, float* bb_sum, void* stream) {
  if (!W_raw || !bb || !match || !n_gt || !partial || !bb_sum || B <= 0 || N <= 0 || K <= 0) return P2C_EINVAL;
  const int KP = kp_of(K);
  if (!KP) return P2C_EUNSUPPORTED;
  const int nchunks = p2c_ceil_div(N, SEG_CHUNK);
  if (partial_elems < (int64_t)B * nchunks) return P2C_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(nchunks, B);
#define P2C_BB_LAUNCH(KPV) \
  bb_partial_kernel<KPV><<<grid, SEG_THREADS, 0, st>>>(W_raw, ldw, bb, match, n_gt, N, K, partial, nchunks)
  if (KP == 2) P2C_BB_LAUNCH(2); else if (KP == 4) P2C_BB_LAUNCH(4);
  else if (KP == 8) P2C_BB_LAUNCH(8); else P2C_BB_LAUNCH(16);
#undef P2C_BB_LAUNCH
  P2C_RETURN_IF_CUDA_ERROR();
  bb_reduce_kernel<<<p2c_ceil_div(B, 128), 128, 0, st>>>(partial, nchunks, B, bb_sum);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_loss_finalize(const float* stats, const float* bb_sum, const int64_t* match,
                                 const int32_t* n_gt, const float* gt_axes, const float* gt_centers,
                                 int B, int N, int K, int norm_eig, const float* weights,
                                 float* E_AX, float* centers, float* per_seg, float* per_cloud,
                                 float* out_losses, void* stream) {
  if (!stats || !match || !n_gt || !gt_axes || !gt_centers || !weights || !E_AX || !centers ||
      !per_seg || !per_cloud || !out_losses || B <= 0 || N <= 0 || K <= 0)
    return P2C_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  loss_segments_kernel<<<p2c_ceil_div(B * K, 4), 128, 0, st>>>(stats, match, n_gt, gt_axes, gt_centers, B,
                                                              N, K, norm_eig, E_AX, centers, per_seg);
  P2C_RETURN_IF_CUDA_ERROR();
  LossWeights lw;
  for (int i = 0; i < 5; ++i) lw.w[i] = weights[i];
  loss_reduce_kernel<<<1, 128, 0, st>>>(stats, bb_sum, n_gt, per_seg, B, N, K, lw, per_cloud, out_losses);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_eig3x3_smallest(const float* M, int n, float* vec, float* eval, void* stream) {
  if (!M || !vec || n <= 0) return P2C_EINVAL;
  eig3x3_kernel<<<p2c_ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(M, n, vec, eval);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
