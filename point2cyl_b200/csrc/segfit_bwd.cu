// Backward of the loss block: gradients of the Point2Cyl losses
// (train_Point2Cyl_without_sketch.py:246-353, losses.py:90-143, data_utils.py:99-177, :253-266) with respect
// to the network outputs, as two kernels that mirror the forward's two levels:
//
//   p2c_loss_backward_coef   per (cloud, gt slot): d total / d statistics  — relaxed IoU quotient rule,
//                            centre difference, and the eigenvector sensitivity of the 3x3 axis fit
//                            (the formula torch.symeig's backward uses: G = sym(sum_i c_i v_i v_0^T),
//                            c_i = <dL/da, v_i> / (lambda_0 - lambda_i)), all in float64;
//   p2c_segfit_backward      one sweep over the points (same KP-lanes-per-point layout as the forward pass):
//                            chain rule from the statistics' gradients through W = barrel + base, the squared
//                            weights of the scatter matrices, the softmax over 2K logits, F.normalize, and the
//                            base/barrel cross-entropy term, writing dW_raw (B,N,2K) and dX_raw (B,N,3).
//
// The statistics are sums over points, so d stat / d input is local to a point: the backward never needs more
// than the (B, stride) gradient row — exactly like the forward never needs more than the statistics row.
// p2c_segfit_backward_w is the same sweep for the function-level API (soft assignments given, no softmax / bb
// term); p2c_eig3x3_backward is the stand-alone eigenvector sensitivity for estimate_extrusion_axis.
#include "segfit_common.cuh"

namespace {

// d L / d m6 (xx xy xz yy yz zz) of a symmetric matrix M given g = d L / d a for its unit eigenvector a of the
// smallest eigenvalue.  Off-diagonal entries of m6 appear twice in M, hence the factor 2.
__device__ void eigvec_sensitivity(const double e[3], const double v[3][3], const double g[3], double dm6[6]) {
  double G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 1; i < 3; ++i) {
    const double gap = e[0] - e[i];
    if (fabs(gap) < 1e-30) continue;  // degenerate pair: torch would return inf; we drop the term
    const double c = (g[0] * v[i][0] + g[1] * v[i][1] + g[2] * v[i][2]) / gap;
    for (int r = 0; r < 3; ++r)
      for (int s = 0; s < 3; ++s) G[r][s] += 0.5 * c * (v[i][r] * v[0][s] + v[0][r] * v[i][s]);
  }
  dm6[0] = G[0][0]; dm6[1] = 2.0 * G[0][1]; dm6[2] = 2.0 * G[0][2];
  dm6[3] = G[1][1]; dm6[4] = 2.0 * G[1][2]; dm6[5] = G[2][2];
}

// eff: effective weights {seg, normal, bb, axis, centre} on the device (already multiplied by the upstream grad)
__global__ void loss_bwd_coef_kernel(const float* __restrict__ stats, const int64_t* __restrict__ match,
                                     const int32_t* __restrict__ n_gt, const float* __restrict__ gt_axes,
                                     const float* __restrict__ gt_centers, const float* __restrict__ eff, int B,
                                     int N, int K, int norm_eig, float* __restrict__ dstats) {
  const int seg = blockIdx.x * blockDim.x + threadIdx.x;
  if (seg >= B * K) return;
  const int b = seg / K, j = seg % K;
  const int stride = seg_stride(K);
  const float* s = stats + (size_t)b * stride;
  float* d = dstats + (size_t)b * stride;
  const int ng = min(n_gt[b], K);
  if (j == 0) d[off_normal(K)] = eff[1] / ((float)B * (float)N);
  if (j >= ng) return;
  const int m = (int)match[seg];
  const double inv_bn = 1.0 / ((double)B * (double)ng);
  // relaxed IoU (losses.py:95-101): L = cs * (1 - D / den), den = cnt + colsum - D + eps
  {
    const double cs = (double)eff[0] * inv_bn;
    const double D = s[j * K + m];
    const double den = (double)s[off_cnt(K) + j] + (double)s[off_colsum(K) + m] - D + 1e-10;
    d[j * K + m] = (float)(-cs * (den + D) / (den * den));
    d[off_colsum(K) + m] = (float)(cs * D / (den * den));
  }
  // centre (data_utils.py:253-266): L = cc * |C/N - c_gt|^2
  {
    const double cc = (double)eff[4] * inv_bn;
    for (int i = 0; i < 3; ++i) {
      const double c = (double)s[off_C(K) + m * 3 + i] / (double)N;
      d[off_C(K) + m * 3 + i] = (float)(cc * 2.0 * (c - (double)gt_centers[(size_t)seg * 3 + i]) / (double)N);
    }
  }
  // axis (data_utils.py:162-172 + losses.py:130): L = ca * (1 - |<a, a_gt>|)
  {
    const double ca = (double)eff[3] * inv_bn;
    double sb = 1.0, sc = 1.0;
    if (norm_eig) {
      const double nb = sqrt((double)s[off_cbar(K) + j]) + 1.0, nc = sqrt((double)s[off_cbase(K) + j]) + 1.0;
      sb = 1.0 / (nb * nb); sc = 1.0 / (nc * nc);
    }
    double a6[6];
    for (int i = 0; i < 6; ++i)
      a6[i] = (double)s[off_Mbar(K) + m * 6 + i] * sb - (double)s[off_Mbase(K) + m * 6 + i] * sc;
    double e[3], v[3][3];
    jacobi3(a6, e, v);
    const float* ga = gt_axes + (size_t)seg * 3;
    const double dot = (double)(float)v[0][0] * ga[0] + (double)(float)v[0][1] * ga[1] + (double)(float)v[0][2] * ga[2];
    const double sg = dot > 0.0 ? 1.0 : (dot < 0.0 ? -1.0 : 0.0);
    const double g[3] = {-ca * sg * ga[0], -ca * sg * ga[1], -ca * sg * ga[2]};
    double dm6[6];
    eigvec_sensitivity(e, v, g, dm6);
    for (int i = 0; i < 6; ++i) {
      d[off_Mbar(K) + m * 6 + i] = (float)(dm6[i] * sb);
      d[off_Mbase(K) + m * 6 + i] = (float)(-dm6[i] * sc);
    }
  }
}

__global__ void eig3x3_bwd_kernel(const float* __restrict__ M, const float* __restrict__ gvec, int n,
                                  float* __restrict__ dM) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* m = M + (size_t)i * 9;
  const double a[6] = {m[0], m[1], m[2], m[4], m[5], m[8]};
  double e[3], v[3][3];
  jacobi3(a, e, v);
  const double g[3] = {gvec[i * 3], gvec[i * 3 + 1], gvec[i * 3 + 2]};
  double dm6[6];
  eigvec_sensitivity(e, v, g, dm6);
  // symmetric gradient over the full 3x3 (torch.symeig's backward symmetrises too)
  float* o = dM + (size_t)i * 9;
  o[0] = (float)dm6[0]; o[4] = (float)dm6[3]; o[8] = (float)dm6[5];
  o[1] = o[3] = (float)(0.5 * dm6[1]); o[2] = o[6] = (float)(0.5 * dm6[2]); o[5] = o[7] = (float)(0.5 * dm6[4]);
}

// LOGITS: W_raw holds 2K logits per point (softmax + bb term inside); otherwise wb/wc are the soft assignments.
template <int KP, bool LOGITS>
__global__ void __launch_bounds__(SEG_THREADS)
segfit_bwd_kernel(const float* __restrict__ X, int64_t ldx, int normalize_x, const float* __restrict__ wbp,
                  int64_t ldb, int64_t sb, const float* __restrict__ wcp, int64_t ldc, int64_t sc,
                  const float* __restrict__ pcs, const float* __restrict__ gtn, const int64_t* __restrict__ inst,
                  const int64_t* __restrict__ bb, const float* __restrict__ dstats,
                  const int64_t* __restrict__ match, const int32_t* __restrict__ n_gt,
                  const float* __restrict__ eff, int B, int N, int K, float* __restrict__ dX, int64_t lddx,
                  float* __restrict__ dWb, int64_t lddb, int64_t sdb, float* __restrict__ dWc, int64_t lddc,
                  int64_t sdc) {
  constexpr int PPS = SEG_THREADS / KP;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tid = threadIdx.x;
  const int k = tid % KP, pl = tid / KP;
  const int lane = tid & 31;
  const int gbase = lane & ~(KP - 1);
  const int stride = seg_stride(K);
  const float* d = dstats + (size_t)b * stride;
  __shared__ float s_dD[16 * 16];
  for (int i = tid; i < K * K; i += SEG_THREADS) s_dD[i] = d[i];
  // per-column gradient coefficients live in registers of the lane that owns the column
  float dcol = 0.f, dC0 = 0.f, dC1 = 0.f, dC2 = 0.f, dmb[6] = {0, 0, 0, 0, 0, 0}, dmc[6] = {0, 0, 0, 0, 0, 0};
  if (k < K) {
    dcol = d[off_colsum(K) + k];
    dC0 = d[off_C(K) + k * 3]; dC1 = d[off_C(K) + k * 3 + 1]; dC2 = d[off_C(K) + k * 3 + 2];
#pragma unroll
    for (int i = 0; i < 6; ++i) { dmb[i] = d[off_Mbar(K) + k * 6 + i]; dmc[i] = d[off_Mbase(K) + k * 6 + i]; }
  }
  const float dnormal = d[off_normal(K)];
  // bb term: slot k takes column my_col; inv_slot = live slot whose matched column is k (or -1)
  int my_col = 0, inv_slot = -1;
  bool slot_live = false;
  float kappa = 0.f;
  if (LOGITS && match) {
    const int ng = min(n_gt[b], K);
    if (k < K) my_col = (int)match[(size_t)b * K + k];
    slot_live = k < K && k < ng;
    for (int j = 0; j < ng; ++j)
      if ((int)match[(size_t)b * K + j] == k) inv_slot = j;
    kappa = eff[2] / ((float)B * (float)N);
  }
  __syncthreads();

  const int n_end = min(N, (chunk + 1) * SEG_CHUNK);
  for (int n0 = chunk * SEG_CHUNK; n0 < n_end; n0 += PPS) {
    const int n = n0 + pl;
    const bool ok = n < n_end;
    const size_t row = (size_t)b * N + (ok ? n : (n_end - 1));
    float rbar = 0.f, rbase = 0.f, wb = 0.f, wc = 0.f;
    if (LOGITS) {
      point_softmax<KP>(wbp + row * ldb, k, K, rbar, rbase, wb, wc);
    } else if (k < K) {
      wb = __ldg(wbp + row * ldb + (size_t)k * sb);
      if (wcp) wc = __ldg(wcp + row * ldc + (size_t)k * sc);
    }
    float x = 0.f, y = 0.f, z = 0.f, inv = 1.f;
    if (X) {
      const float* xr = X + row * ldx;
      x = __ldg(xr); y = __ldg(xr + 1); z = __ldg(xr + 2);
      if (normalize_x) {
        inv = 1.0f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
        x *= inv; y *= inv; z *= inv;
      }
    }
    const int g = inst ? (int)inst[row] : -1;
    // gradient w.r.t. W[k] = wb + wc (shared by both halves)
    float gW = dcol;
    if (g >= 0 && g < K && k < K) gW += s_dD[g * K + k];
    if (pcs) {
      const float* pr = pcs + row * 3;
      gW += dC0 * __ldg(pr) + dC1 * __ldg(pr + 1) + dC2 * __ldg(pr + 2);
    }
    const float xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
    const float qb = dmb[0] * xx + dmb[1] * xy + dmb[2] * xz + dmb[3] * yy + dmb[4] * yz + dmb[5] * zz;
    const float qc = dmc[0] * xx + dmc[1] * xy + dmc[2] * xz + dmc[3] * yy + dmc[4] * yz + dmc[5] * zz;
    float gbar = gW + 2.f * wb * qb;
    float gbas = gW + 2.f * wc * qc;
    // d / d x_hat from the scatter matrices: T_i = sum_k wb^2 dMbar[k][i] + wc^2 dMbase[k][i]
    float T[6];
    const float b2 = wb * wb, c2 = wc * wc;
#pragma unroll
    for (int i = 0; i < 6; ++i) T[i] = group_sum<KP>(b2 * dmb[i] + c2 * dmc[i]);
    float gx = 2.f * T[0] * x + T[1] * y + T[2] * z;
    float gy = T[1] * x + 2.f * T[3] * y + T[4] * z;
    float gz = T[2] * x + T[4] * y + 2.f * T[5] * z;
    if (gtn && X) {
      const float* gr = gtn + row * 3;
      const float g0 = __ldg(gr), g1 = __ldg(gr + 1), g2 = __ldg(gr + 2);
      const float dot = x * g0 + y * g1 + z * g2;
      const float sg = dot > 0.f ? 1.f : (dot < 0.f ? -1.f : 0.f);
      gx -= dnormal * sg * g0; gy -= dnormal * sg * g1; gz -= dnormal * sg * g2;
    }
    float ce_bar = 0.f, ce_base = 0.f;
    if (LOGITS && match) {
      // forward of the bb term (train_...:286-307) recomputed, then its two gradient paths
      const float w = wb + wc;
      const float wm = __shfl_sync(P2C_FULL_MASK, w, gbase + my_col);
      const float zin = slot_live ? wm : 0.f;
      const float NEG = -__int_as_float(0x7f800000);
      const float zmax = group_max<KP>(k < K ? zin : NEG);
      const float ez = k < K ? expf(zin - zmax) : 0.f;
      const float Z = ez / group_sum<KP>(ez);
      float ce = 0.f, pbar = 0.f, pbase = 0.f;
      const int t = (int)bb[row];
      if (k < K) {
        const float mx = fmaxf(rbar, rbase);
        const float eb = expf(rbar - mx), ec = expf(rbase - mx);
        const float lse = mx + logf(eb + ec);
        ce = lse - (t == 0 ? rbar : rbase);
        pbar = eb / (eb + ec); pbase = ec / (eb + ec);
      }
      const float mean_ce = group_sum<KP>(Z * ce);
      const float du = slot_live ? kappa * Z * (ce - mean_ce) : 0.f;   // d / d (matched W of slot k)
      const float du_col = __shfl_sync(P2C_FULL_MASK, du, gbase + (inv_slot >= 0 ? inv_slot : 0));
      if (inv_slot >= 0) { gbar += du_col; gbas += du_col; }
      ce_bar = kappa * Z * (pbar - (t == 0 ? 1.f : 0.f));
      ce_base = kappa * Z * (pbase - (t == 0 ? 0.f : 1.f));
    }
    float out_b = gbar, out_c = gbas;
    if (LOGITS) {
      const float dotp = group_sum<KP>(wb * gbar + wc * gbas);
      out_b = wb * (gbar - dotp) + ce_bar;
      out_c = wc * (gbas - dotp) + ce_base;
    }
    if (ok && k < K) {
      if (dWb) dWb[row * lddb + (size_t)(LOGITS ? 2 * k : k) * sdb] = out_b;
      if (LOGITS) dWb[row * lddb + (size_t)(2 * k + 1) * sdb] = out_c;
      else if (dWc) dWc[row * lddc + (size_t)k * sdc] = out_c;
    }
    if (ok && k == 0 && dX) {
      if (normalize_x) {
        const float pr = gx * x + gy * y + gz * z;
        gx = (gx - pr * x) * inv; gy = (gy - pr * y) * inv; gz = (gz - pr * z) * inv;
      }
      float* o = dX + row * lddx;
      o[0] = gx; o[1] = gy; o[2] = gz;
    }
  }
}

}  // namespace

extern "C" int p2c_loss_backward_coef(const float* stats, const int64_t* match, const int32_t* n_gt,
                                      const float* gt_axes, const float* gt_centers, const float* eff_weights,
                                      int B, int N, int K, int norm_eig, float* dstats, void* stream) {
  if (!stats || !match || !n_gt || !gt_axes || !gt_centers || !eff_weights || !dstats || B <= 0 || N <= 0 || K <= 0)
    return P2C_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  P2C_CUDA_TRY(cudaMemsetAsync(dstats, 0, sizeof(float) * (size_t)B * seg_stride(K), st));
  loss_bwd_coef_kernel<<<p2c_ceil_div(B * K, 64), 64, 0, st>>>(stats, match, n_gt, gt_axes, gt_centers, eff_weights, B,
                                                              N, K, norm_eig, dstats);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

#define P2C_SEGB_LAUNCH(KPV, LG)                                                                                   \
  segfit_bwd_kernel<KPV, LG><<<grid, SEG_THREADS, 0, st>>>(X, ldx, normalize_x, wb, ldb, sb, wc, ldc, sc, pcs,      \
                                                           gt_normals, inst, bb, dstats, match, n_gt, eff, B, N, K, \
                                                           dX, lddx, dWb, lddb, sdb, dWc, lddc, sdc)

extern "C" int p2c_segfit_backward(const float* X_raw, int64_t ldx, const float* W_raw, int64_t ldw, const float* pcs,
                                   const float* gt_normals, const int64_t* inst, const int64_t* bb,
                                   const float* dstats, const int64_t* match, const int32_t* n_gt,
                                   const float* eff, int B, int N, int K, float* dX_raw, int64_t lddx,
                                   float* dW_raw, int64_t lddw, void* stream) {
  if (!X_raw || !W_raw || !pcs || !gt_normals || !inst || !bb || !dstats || !dX_raw || !dW_raw) return P2C_EINVAL;
  if ((match == nullptr) != (n_gt == nullptr) || (match && !eff)) return P2C_EINVAL;
  if (B <= 0 || N <= 0 || K <= 0 || ldx < 3 || ldw < 2 * K || lddx < 3 || lddw < 2 * K) return P2C_EINVAL;
  const int KP = kp_of(K);
  if (!KP) return P2C_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(p2c_ceil_div(N, SEG_CHUNK), B);
  const float* X = X_raw; const int normalize_x = 1;
  const float* wb = W_raw; const int64_t ldb = ldw, sb = 1; const float* wc = nullptr; const int64_t ldc = 0, sc = 0;
  float* dX = dX_raw; float* dWb = dW_raw; const int64_t lddb = lddw, sdb = 1; float* dWc = nullptr;
  const int64_t lddc = 0, sdc = 0;
  if (KP == 2) P2C_SEGB_LAUNCH(2, true); else if (KP == 4) P2C_SEGB_LAUNCH(4, true);
  else if (KP == 8) P2C_SEGB_LAUNCH(8, true); else P2C_SEGB_LAUNCH(16, true);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_segfit_backward_w(const float* X, int64_t ldx, int normalize_x, const float* wb, int64_t ldb,
                                     int64_t sb, const float* wc, int64_t ldc, int64_t sc, const float* pcs,
                                     const float* gt_normals, const int64_t* inst, const float* dstats, int B, int N,
                                     int K, float* dX, int64_t lddx, float* dWb, float* dWc, void* stream) {
  if (!wb || !dstats || B <= 0 || N <= 0 || K <= 0) return P2C_EINVAL;
  if (dX && (!X || lddx < 3)) return P2C_EINVAL;
  if (dWc && !wc) return P2C_EINVAL;
  const int KP = kp_of(K);
  if (!KP) return P2C_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(p2c_ceil_div(N, SEG_CHUNK), B);
  const int64_t* bb = nullptr; const int64_t* match = nullptr; const int32_t* n_gt = nullptr; const float* eff = nullptr;
  const int64_t lddb = K, sdb = 1, lddc = K, sdc = 1;   // gradients are written contiguous (B,N,K)
  if (KP == 2) P2C_SEGB_LAUNCH(2, false); else if (KP == 4) P2C_SEGB_LAUNCH(4, false);
  else if (KP == 8) P2C_SEGB_LAUNCH(8, false); else P2C_SEGB_LAUNCH(16, false);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
#undef P2C_SEGB_LAUNCH

extern "C" int p2c_eig3x3_backward(const float* M, const float* gvec, int n, float* dM, void* stream) {
  if (!M || !gvec || !dM || n <= 0) return P2C_EINVAL;
  eig3x3_bwd_kernel<<<p2c_ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(M, gvec, n, dM);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
