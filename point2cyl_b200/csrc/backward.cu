// Backward of the PointNet++ backbone: what autograd runs for
// models/pointnet_util.py:181-207 (set abstraction), :283-320 (feature propagation) and
// models/pointnet_extrusion.py:58-65 (head) in the reference, restated on the forward's own data layout:
// activations are raw (pre-BatchNorm) row matrices, BatchNorm+ReLU of layer i lives in layer i+1's operand load.
//
// Per MLP layer, given dA = d L / d relu(bn(Y)):
//   p2c_bn_bwd_reduce   s1 = sum_m dYh, s2 = sum_m dYh * Y with dYh = dA * [scale*Y + shift > 0]      (1 pass)
//   p2c_bn_bwd_coef     dgamma, dbeta and the per-channel affine (a, b, c) of BatchNorm's backward:
//                       dY = a*dYh + b*Y + c   (train: a = gamma*invstd, b = -a*invstd*dgamma/M,
//                       c = -a*dbeta/M - b*mean; eval: b = c = 0)
//   p2c_bn_bwd_apply    dY materialised once (in place over dA)
//   p2c_linear          dA_prev = dY * W   (the forward's tcgen05 kernel on the transposed weight)
//   p2c_wgrad           dW += dY^T * A_prev with A_prev = relu(bn(X_prev)) recomputed in the operand load, db += sum dY
// Max-pooled layers (last layer of a set-abstraction level): the gradient enters only at the arg-max row of each
// group, recovered as the first row whose raw value equals the pooled max (min when scale < 0) — the forward
// stores no indices (p2c_pool_bwd_reduce / p2c_pool_bwd_apply).
// The fused gather + first SA conv runs backward as a scatter-add into the per-source-point product Qf
// (p2c_sa_first_bwd), the 3-NN interpolation as a weighted scatter-add (p2c_three_nn_interp_bwd), the heads as
// p2c_head_bwd (data gradient + re-materialised masked input) followed by p2c_wgrad on plain row matrices.
// p2c_wgrad dispatches to the tcgen05 kernel of wgrad_tc.cu; the SIMT kernel below takes masked / unaligned / tiny
// calls and P2C_PREC_FP32.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------
// BatchNorm + ReLU backward reductions
// ---------------------------------------------------------------------------------------------------------

constexpr int RED_THREADS = 256;
constexpr int RED_CB = 128;     // channels per CTA (32 float4 lanes); C < 128 uses C/4 lanes and more row lanes

// CTA = (row slab blockIdx.x, channel block blockIdx.y); thread = (float4 channel lane, row lane); float4 loads
__global__ void __launch_bounds__(RED_THREADS)
bn_bwd_reduce_kernel(const float* __restrict__ dA, int64_t ldda, const float* __restrict__ Y, int64_t ldy,
                     const float* __restrict__ scale, const float* __restrict__ shift, int64_t M, int C,
                     int rows_per_cta, double* __restrict__ sums) {
  __shared__ float s_red[RED_THREADS][8];
  const int cb = C < RED_CB ? C : RED_CB;            // channels of this CTA's block
  const int CL = cb >> 2;                            // float4 lanes (8, 16 or 32)
  const int RL = RED_THREADS / CL;                   // row lanes
  const int cl = threadIdx.x % CL, rl = threadIdx.x / CL;
  const int c = blockIdx.y * cb + cl * 4;
  // slabs are walked from the LAST rows to the first: dA has just been written front to back by the data-gradient
  // kernel, so its tail is what the L2 still holds - and the applying kernel that follows starts at row 0, where this
  // kernel ends
  const int64_t r0 = (int64_t)(gridDim.x - 1 - blockIdx.x) * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (scale) {
    sc = __ldg(reinterpret_cast<const float4*>(scale + c));
    sh = __ldg(reinterpret_cast<const float4*>(shift + c));
  }
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  auto accumulate = [&](const float4& y, float4 g) {
    if (scale) {
      if (!(fmaf(y.x, sc.x, sh.x) > 0.f)) g.x = 0.f;
      if (!(fmaf(y.y, sc.y, sh.y) > 0.f)) g.y = 0.f;
      if (!(fmaf(y.z, sc.z, sh.z) > 0.f)) g.z = 0.f;
      if (!(fmaf(y.w, sc.w, sh.w) > 0.f)) g.w = 0.f;
    }
    s1.x += g.x; s1.y += g.y; s1.z += g.z; s1.w += g.w;
    s2.x = fmaf(g.x, y.x, s2.x); s2.y = fmaf(g.y, y.y, s2.y); s2.z = fmaf(g.z, y.z, s2.z); s2.w = fmaf(g.w, y.w, s2.w);
  };
  // four rows (eight 16-byte loads) in flight per thread before the arithmetic starts
  int64_t r = r0 + rl;
  for (; r + 3 * RL < r1; r += 4 * RL) {
    float4 y[4], g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      y[u] = __ldg(reinterpret_cast<const float4*>(Y + (r + u * RL) * ldy + c));
      g[u] = __ldg(reinterpret_cast<const float4*>(dA + (r + u * RL) * ldda + c));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) accumulate(y[u], g[u]);
  }
  for (; r < r1; r += RL)
    accumulate(__ldg(reinterpret_cast<const float4*>(Y + r * ldy + c)), __ldg(reinterpret_cast<const float4*>(dA + r * ldda + c)));
  float* mine = s_red[threadIdx.x];
  mine[0] = s1.x; mine[1] = s1.y; mine[2] = s1.z; mine[3] = s1.w;
  mine[4] = s2.x; mine[5] = s2.y; mine[6] = s2.z; mine[7] = s2.w;
  __syncthreads();
  // thread t < 8*CL: (lane cl', component q) summed over the row lanes
  for (int e = threadIdx.x; e < 8 * CL; e += RED_THREADS) {
    const int l = e >> 3, q = e & 7;
    float t = 0.f;
    for (int j = 0; j < RL; ++j) t += s_red[j * CL + l][q];
    const int ch = blockIdx.y * cb + l * 4 + (q & 3);
    atomicAdd(sums + (q < 4 ? 0 : C) + ch, (double)t);
  }
}

__global__ void __launch_bounds__(256)
pool_bwd_reduce_kernel(const float* __restrict__ dOut, int64_t ldd, const float* __restrict__ Ymax,
                       const float* __restrict__ Ymin, const float* __restrict__ scale,
                       const float* __restrict__ shift, int64_t G, int C, double* __restrict__ sums) {
  // one thread per channel (grid.x over channel blocks), grid.y over group slabs
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int64_t g0 = (int64_t)blockIdx.y * 64;
  const int64_t g1 = g0 + 64 < G ? g0 + 64 : G;
  const float sc = __ldg(scale + c), sh = __ldg(shift + c);
  float s1 = 0.f, s2 = 0.f;
  for (int64_t g = g0; g < g1; ++g) {
    const float y = sc >= 0.f ? __ldg(Ymax + g * C + c) : __ldg(Ymin + g * C + c);
    float d = __ldg(dOut + g * ldd + c);
    if (!(fmaf(y, sc, sh) > 0.f)) d = 0.f;
    s1 += d;
    s2 = fmaf(d, y, s2);
  }
  atomicAdd(sums + c, (double)s1);
  atomicAdd(sums + C + c, (double)s2);
}

// The per-channel part of BatchNorm's backward, evaluated by the kernel that APPLIES it (p2c_bn_bwd_apply_fused /
// p2c_pool_bwd_apply_fused) instead of by a micro-launch per layer: same float64 arithmetic as bn_bwd_coef_kernel.
struct BnBwdCoefIn {
  const double* sums; double count;
  const float* gamma; const float* mean; const float* invstd;
  int training;
  float* dgamma; float* dbeta;
};
__device__ __forceinline__ void bn_bwd_coef_channel(const BnBwdCoefIn& ci, int c, int C, bool writer, float& a_out,
                                                    float& b_out, float& c_out) {
  const double s1 = ci.sums[c], s2 = ci.sums[C + c];
  const double mu = ci.mean[c], is = ci.invstd[c], g = ci.gamma ? (double)ci.gamma[c] : 1.0;
  const double dg = is * (s2 - mu * s1);
  const double a = g * is;
  double b = 0.0, cc = 0.0;
  if (ci.training) {
    b = -a * is * dg / ci.count;
    cc = -a * s1 / ci.count - b * mu;
  }
  a_out = (float)a; b_out = (float)b; c_out = (float)cc;
  if (writer) {
    if (ci.dgamma) ci.dgamma[c] += (float)dg;
    if (ci.dbeta) ci.dbeta[c] += (float)s1;
  }
}

__global__ void bn_bwd_coef_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ mean, const float* __restrict__ invstd, int training,
                                   float* __restrict__ coef, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                   int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double s1 = sums[c], s2 = sums[C + c];
  const double mu = mean[c], is = invstd[c], g = gamma ? (double)gamma[c] : 1.0;
  const double dg = is * (s2 - mu * s1);
  const double a = g * is;
  double b = 0.0, cc = 0.0;
  if (training) {
    b = -a * is * dg / count;
    cc = -a * s1 / count - b * mu;
  }
  coef[c] = (float)a;
  coef[C + c] = (float)b;
  coef[2 * C + c] = (float)cc;
  if (dgamma) dgamma[c] += (float)dg;
  if (dbeta) dbeta[c] += (float)s1;
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ dA, int64_t ldda, const float* __restrict__ Y, int64_t ldy,
                    const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ coef, int64_t M, int C, float* __restrict__ dY, int64_t lddy) {
  const int C4 = C >> 2;
  const int64_t total = M * C4;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = e / C4;
    const int c = (int)(e - m * C4) * 4;
    const float4 y = __ldg(reinterpret_cast<const float4*>(Y + m * ldy + c));
    float4 g = __ldg(reinterpret_cast<const float4*>(dA + m * ldda + c));
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c));
    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + c));
    const float4 a = __ldg(reinterpret_cast<const float4*>(coef + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(coef + C + c));
    const float4 k = __ldg(reinterpret_cast<const float4*>(coef + 2 * C + c));
    if (!(fmaf(y.x, sc.x, sh.x) > 0.f)) g.x = 0.f;
    if (!(fmaf(y.y, sc.y, sh.y) > 0.f)) g.y = 0.f;
    if (!(fmaf(y.z, sc.z, sh.z) > 0.f)) g.z = 0.f;
    if (!(fmaf(y.w, sc.w, sh.w) > 0.f)) g.w = 0.f;
    float4 o;
    o.x = fmaf(a.x, g.x, fmaf(b.x, y.x, k.x));
    o.y = fmaf(a.y, g.y, fmaf(b.y, y.y, k.y));
    o.z = fmaf(a.z, g.z, fmaf(b.z, y.z, k.z));
    o.w = fmaf(a.w, g.w, fmaf(b.w, y.w, k.w));
    *reinterpret_cast<float4*>(dY + m * lddy + c) = o;
  }
}

// p2c_bn_bwd_apply with the coefficients computed in the kernel: every CTA evaluates the C channels once (float64,
// shared memory); the launch geometry keeps a thread on ONE group of four channels (gridDim.x * 256 is a multiple of
// C / 4), so a, b, c are twelve registers read once per thread
__global__ void __launch_bounds__(256)
bn_bwd_apply_fused_kernel(const float* __restrict__ dA, int64_t ldda, const float* __restrict__ Y, int64_t ldy,
                          const float* __restrict__ scale, const float* __restrict__ shift, const BnBwdCoefIn ci,
                          int64_t M, int C, float* __restrict__ dY, int64_t lddy) {
  extern __shared__ __align__(16) float s_coef[];      // [3][C]: a | b | c, one float64 evaluation per channel and CTA
  const int C4 = C >> 2;
  const int64_t total = M * C4;
  const int64_t e0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = (int)(e0 % C4) * 4;
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x)
    bn_bwd_coef_channel(ci, ch, C, blockIdx.x == 0, s_coef[ch], s_coef[C + ch], s_coef[2 * C + ch]);
  __syncthreads();
  float a[4], b[4], k[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { a[i] = s_coef[c + i]; b[i] = s_coef[C + c + i]; k[i] = s_coef[2 * C + c + i]; }
  const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c));
  const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + c));
  for (int64_t e = e0; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = e / C4;
    const float4 y = __ldg(reinterpret_cast<const float4*>(Y + m * ldy + c));
    float4 g = __ldg(reinterpret_cast<const float4*>(dA + m * ldda + c));
    if (!(fmaf(y.x, sc.x, sh.x) > 0.f)) g.x = 0.f;
    if (!(fmaf(y.y, sc.y, sh.y) > 0.f)) g.y = 0.f;
    if (!(fmaf(y.z, sc.z, sh.z) > 0.f)) g.z = 0.f;
    if (!(fmaf(y.w, sc.w, sh.w) > 0.f)) g.w = 0.f;
    float4 o;
    o.x = fmaf(a[0], g.x, fmaf(b[0], y.x, k[0]));
    o.y = fmaf(a[1], g.y, fmaf(b[1], y.y, k[1]));
    o.z = fmaf(a[2], g.z, fmaf(b[2], y.z, k[2]));
    o.w = fmaf(a[3], g.w, fmaf(b[3], y.w, k[3]));
    *reinterpret_cast<float4*>(dY + m * lddy + c) = o;
  }
}

// one thread per (group, channel): walk the group's rows, route dOut to the first row that attains the pooled value
template <bool FUSED>
__global__ void __launch_bounds__(256)
pool_bwd_apply_kernel(const float* __restrict__ dOut, int64_t ldd, const float* __restrict__ Ymax,
                      const float* __restrict__ Ymin, const float* __restrict__ Y, int64_t ldy,
                      const float* __restrict__ scale, const float* __restrict__ shift,
                      const float* __restrict__ coef, const BnBwdCoefIn ci, int64_t G, int group, int C,
                      float* __restrict__ dY, int64_t lddy) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= G * C) return;
  const int64_t g = e / C;
  const int c = (int)(e - g * C);
  const float sc = __ldg(scale + c), sh = __ldg(shift + c);
  float a, b, k;
  if (FUSED) bn_bwd_coef_channel(ci, c, C, g == 0, a, b, k);
  else { a = __ldg(coef + c); b = __ldg(coef + C + c); k = __ldg(coef + 2 * C + c); }
  const float ysel = sc >= 0.f ? __ldg(Ymax + g * C + c) : __ldg(Ymin + g * C + c);
  float d = __ldg(dOut + g * ldd + c);
  if (!(fmaf(ysel, sc, sh) > 0.f)) d = 0.f;
  bool hit = false;
  const int64_t r0 = g * group;
  for (int j = 0; j < group; ++j) {
    const float y = __ldg(Y + (r0 + j) * ldy + c);
    float o = fmaf(b, y, k);
    if (!hit && y == ysel) { o = fmaf(a, d, o); hit = true; }
    dY[(r0 + j) * lddy + c] = o;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Weight gradient: dW[n,k] += sum_m dY[m,n] * A[m,k],  A = f(X) (BN+ReLU fold, optional channel-first mask)
// SIMT register-tiled, the row dimension split over the grid, fp32 atomics into dW.
// ---------------------------------------------------------------------------------------------------------

constexpr int WG_RS = 16;   // rows per shared-memory stage

template <int TN, int TK>
__global__ void __launch_bounds__(256)
wgrad_kernel(const float* __restrict__ dY, int64_t lddy, const float* __restrict__ X, int64_t ldx,
             const float* __restrict__ in_scale, const float* __restrict__ in_shift,
             const float* __restrict__ mask_cf, int mask_N, int64_t M, int N, int K, int rows_per_cta,
             float* __restrict__ dW, int64_t lddw, float* __restrict__ db) {
  constexpr int RN = TN / 16, RK = TK / 16;
  __shared__ __align__(16) float s_dy[WG_RS][TN];
  __shared__ __align__(16) float s_a[WG_RS][TK];
  const int n0 = blockIdx.x * TN, k0 = blockIdx.y * TK;
  const int64_t m0 = (int64_t)blockIdx.z * rows_per_cta;
  const int64_t m1 = m0 + rows_per_cta < M ? m0 + rows_per_cta : M;
  const int tid = threadIdx.x;
  const int tn = tid / 16, tk = tid % 16;     // thread owns n = n0 + tn*RN + i, k = k0 + tk*RK + j
  float acc[RN][RK];
#pragma unroll
  for (int i = 0; i < RN; ++i)
#pragma unroll
    for (int j = 0; j < RK; ++j) acc[i][j] = 0.f;
  float bsum[RN];
#pragma unroll
  for (int i = 0; i < RN; ++i) bsum[i] = 0.f;

  for (int64_t ms = m0; ms < m1; ms += WG_RS) {
    // stage WG_RS rows of dY (TN wide) and A (TK wide)
    for (int e = tid; e < WG_RS * TN; e += 256) {
      const int r = e / TN, c = e - r * TN;
      const int64_t m = ms + r;
      s_dy[r][c] = (m < m1 && n0 + c < N) ? __ldg(dY + m * lddy + n0 + c) : 0.f;
    }
    for (int e = tid; e < WG_RS * TK; e += 256) {
      const int r = e / TK, c = e - r * TK;
      const int64_t m = ms + r;
      float v = 0.f;
      if (m < m1 && k0 + c < K) {
        v = __ldg(X + m * ldx + k0 + c);
        if (in_scale) v = fmaxf(fmaf(v, __ldg(in_scale + k0 + c), __ldg(in_shift + k0 + c)), 0.f);
        if (mask_cf) {
          const int64_t b = m / mask_N;
          v *= __ldg(mask_cf + ((size_t)b * K + (k0 + c)) * mask_N + (m - b * mask_N));
        }
      }
      s_a[r][c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < WG_RS; ++r) {
      float dyv[RN], av[RK];
#pragma unroll
      for (int i = 0; i < RN; ++i) dyv[i] = s_dy[r][tn * RN + i];
#pragma unroll
      for (int j = 0; j < RK; ++j) av[j] = s_a[r][tk * RK + j];
#pragma unroll
      for (int i = 0; i < RN; ++i) {
        bsum[i] += dyv[i];
#pragma unroll
        for (int j = 0; j < RK; ++j) acc[i][j] = fmaf(dyv[i], av[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < RN; ++i) {
    const int n = n0 + tn * RN + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < RK; ++j) {
      const int k = k0 + tk * RK + j;
      if (k < K) atomicAdd(dW + (size_t)n * lddw + k, acc[i][j]);
    }
    if (db && blockIdx.y == 0 && tk == 0) atomicAdd(db + n, bsum[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Fused gather + first SA conv, backward: dQf[p] += dY0[r], dWx += dY0^T * (xyz[p] - centre), dbias += sum dY0
// ---------------------------------------------------------------------------------------------------------

template <int CPL>
__global__ void __launch_bounds__(256)
sa_first_bwd_kernel(const float* __restrict__ dY, int64_t lddy, const float* __restrict__ xyz,
                    const float* __restrict__ new_xyz, const int64_t* __restrict__ idx, int N, int S, int ns,
                    int64_t rows, float* __restrict__ dQf, int64_t ldq, float* __restrict__ dW, int64_t lddw,
                    float* __restrict__ dbias) {
  constexpr int C = 32 * CPL;
  __shared__ float s_red[8][4][C];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = lane * CPL;
  float gw[CPL][3], gb[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) { gw[i][0] = gw[i][1] = gw[i][2] = 0.f; gb[i] = 0.f; }
  // 32 rows per warp and step (same scheme as the forward kernel): lane l resolves row r0 + l's index and centred
  // coordinates, the warp then walks the rows with shuffled broadcasts; lane = CPL consecutive channels.
  const int64_t wstride = (int64_t)gridDim.x * 8 * 32;
  for (int64_t r0 = ((int64_t)blockIdx.x * 8 + warp) * 32; r0 < rows; r0 += wstride) {
    const int64_t rl = r0 + lane;
    const bool okl = rl < rows;
    const unsigned rr = (unsigned)(okl ? rl : rows - 1);
    const unsigned bsl = rr / (unsigned)ns, bl = bsl / (unsigned)S;
    int64_t pl = __ldg(idx + rr);
    pl = (pl < 0 || pl >= N) ? 0 : pl;
    const unsigned src = bl * (unsigned)N + (unsigned)pl;
    const float* pp = xyz + (size_t)src * 3;
    const float* cc = new_xyz + (size_t)bsl * 3;
    const float dl0 = __ldg(pp) - __ldg(cc), dl1 = __ldg(pp + 1) - __ldg(cc + 1), dl2 = __ldg(pp + 2) - __ldg(cc + 2);
    const int nrow = (int)min((int64_t)32, rows - r0);
    // eight rows of dY in flight before the first is consumed (ncu: 19 % of the issue slots busy, long-scoreboard
    // stalls - the latency of these loads, not their bandwidth, was the bound: 139 us for 268 MB)
    for (int j0 = 0; j0 < nrow; j0 += 8) {
      float g[8][CPL];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) g[u][i] = 0.f;
        if (j0 + u < nrow) {
          const float* gr = dY + (r0 + j0 + u) * lddy + c0;
          if (CPL == 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(gr));
            g[u][0] = t.x; g[u][1] = t.y; g[u][2] = t.z; g[u][CPL - 1] = t.w;
          } else {
            const float2 t = __ldg(reinterpret_cast<const float2*>(gr));
            g[u][0] = t.x; g[u][1] = t.y;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u;
        if (j >= nrow) break;
        const float d0 = __shfl_sync(P2C_FULL_MASK, dl0, j), d1 = __shfl_sync(P2C_FULL_MASK, dl1, j),
                    d2 = __shfl_sync(P2C_FULL_MASK, dl2, j);
        if (dQf) {
          const unsigned sj = __shfl_sync(P2C_FULL_MASK, src, j);
          float* q = dQf + (size_t)sj * ldq + c0;
          // one 16-byte vector reduction instead of four scalar ones where the row allows it (L2 atomic operations are
          // what this scatter costs)
          if (CPL == 4 && (reinterpret_cast<uintptr_t>(q) & 15) == 0) {
            atomicAdd(reinterpret_cast<float4*>(q), make_float4(g[u][0], g[u][1], g[u][2], g[u][CPL - 1]));
          } else if (CPL == 2 && (reinterpret_cast<uintptr_t>(q) & 7) == 0) {
            atomicAdd(reinterpret_cast<float2*>(q), make_float2(g[u][0], g[u][1]));
          } else {
#pragma unroll
            for (int i = 0; i < CPL; ++i) atomicAdd(q + i, g[u][i]);
          }
        }
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          gw[i][0] = fmaf(g[u][i], d0, gw[i][0]); gw[i][1] = fmaf(g[u][i], d1, gw[i][1]);
          gw[i][2] = fmaf(g[u][i], d2, gw[i][2]);
          gb[i] += g[u][i];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    s_red[warp][0][c0 + i] = gw[i][0]; s_red[warp][1][c0 + i] = gw[i][1]; s_red[warp][2][c0 + i] = gw[i][2];
    s_red[warp][3][c0 + i] = gb[i];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 4 * C; e += 256) {
    const int which = e / C, c = e - which * C;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[w][which][c];
    if (which < 3) atomicAdd(dW + (size_t)c * lddw + which, t);
    else if (dbias) atomicAdd(dbias + c, t);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Grouping gather backward (the un-fused set-abstraction path: first-layer width other than 64 / 128):
// dfeats[b*N + idx[r], :] += dRows[r, 3:3+D]   (the xyz columns carry no gradient: coordinates are data)
// ---------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
group_bwd_kernel(const float* __restrict__ dRows, int64_t ldr, const int64_t* __restrict__ idx, int N, int S, int ns,
                 int D, int64_t rows, float* __restrict__ dF, int64_t ldf) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int64_t b = (r / ns) / S;
  const int64_t src = __ldg(idx + r);
  if (src < 0 || src >= N) return;            // a query without hits gathered zeros in the forward
  const float* g = dRows + r * ldr + 3;
  float* f = dF + ((size_t)b * N + src) * ldf;
  for (int c = lane; c < D; c += 32) atomicAdd(f + c, __ldg(g + c));
}

// ---------------------------------------------------------------------------------------------------------
// 3-NN interpolation backward: dfeats2[b, idx[b,n,j], :] += w[b,n,j] * dInterp[b*N+n, :]
// ---------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
interp_bwd_kernel(const float* __restrict__ dI, int64_t ldi, const int64_t* __restrict__ idx,
                  const float* __restrict__ w, int N, int S, int D, int64_t rows, float* __restrict__ dF,
                  int64_t ldf) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int64_t b = r / N;
  int64_t i0 = __ldg(idx + r * 3), i1 = __ldg(idx + r * 3 + 1), i2 = __ldg(idx + r * 3 + 2);
  const float w0 = __ldg(w + r * 3), w1 = __ldg(w + r * 3 + 1), w2 = __ldg(w + r * 3 + 2);
  float* f0 = dF + ((size_t)b * S + i0) * ldf;
  float* f1 = dF + ((size_t)b * S + i1) * ldf;
  float* f2 = dF + ((size_t)b * S + i2) * ldf;
  const float* g = dI + r * ldi;
  // lanes = consecutive channels: a warp's scalar reduction is ONE 128-byte line per instruction, which the L2 takes as
  // one request.  (Measured: 16-byte vector reductions - 512 bytes per instruction - are slower here, 86 -> 115 us at
  // fp1's 262,144 x 128; they pay where every lane targets a different row, as in the weight-gradient epilogue.)
  for (int c = lane; c < D; c += 32) {
    const float v = __ldg(g + c);
    atomicAdd(f0 + c, w0 * v);
    atomicAdd(f1 + c, w1 * v);
    atomicAdd(f2 + c, w2 * v);
  }
}

// S == 1 (broadcast, models/pointnet_util.py:298-299): dfeats2[b, :] = sum_n dInterp[b*N+n, :]
__global__ void __launch_bounds__(256)
broadcast_bwd_kernel(const float* __restrict__ dI, int64_t ldi, int N, int D, float* __restrict__ dF, int64_t ldf) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  __shared__ float s[8][32];
  float t = 0.f;
  if (c < D)
    for (int n = rl; n < N; n += 8) t += __ldg(dI + ((size_t)b * N + n) * ldi + c);
  s[rl][threadIdx.x & 31] = t;
  __syncthreads();
  if (rl == 0 && c < D) {
    for (int j = 1; j < 8; ++j) t += s[j][threadIdx.x & 31];
    dF[(size_t)b * ldf + c] = t;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Heads backward (data gradient): dA[m,k] = mask[b,k,n] * sum_j dOut[m,j] * W[j,k]
// ---------------------------------------------------------------------------------------------------------

template <int NP>
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ dOut, int64_t ldo, const float* __restrict__ mask_cf,
                const int64_t* __restrict__ seed, const float* __restrict__ W, int64_t M, int N, int C, int Nout, float* __restrict__ dA,
                int64_t ldda, const float* __restrict__ H, int64_t ldh, const float* __restrict__ scale,
                const float* __restrict__ shift, float* __restrict__ A_out, int64_t lda) {
  extern __shared__ __align__(16) float s_w[];   // [C][NP]  (W transposed), then scale[C], shift[C]
  float* s_sc = s_w + C * NP;
  float* s_sh = s_sc + C;
  for (int e = threadIdx.x; e < C * NP; e += 256) {
    const int k = e / NP, j = e - k * NP;
    s_w[e] = j < Nout ? __ldg(W + (size_t)j * C + k) : 0.f;
  }
  for (int k = threadIdx.x; k < C; k += 256) {
    s_sc[k] = scale ? __ldg(scale + k) : 1.f;
    s_sh[k] = shift ? __ldg(shift + k) : 0.f;
  }
  __syncthreads();
  const int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (m >= M) return;
  const int64_t b = m / N;
  const int n = (int)(m - b * N);
  float g[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) g[j] = j < Nout ? __ldg(dOut + m * ldo + j) : 0.f;
  const float* mk = mask_cf ? mask_cf + (size_t)b * C * N + n : nullptr;
  float* out = dA + m * ldda;
  P2CPhilox4 bits;
  for (int k0 = 0; k0 < C; k0 += 4) {
    float v[4], mq[4];
    if (seed && !mk && (k0 & 127) == 0) bits = p2c_dropout_bits(seed, m, k0 >> 7);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float* wr = s_w + (k0 + i) * NP;
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < NP; ++j) t = fmaf(g[j], wr[j], t);
      mq[i] = mk ? __ldg(mk + (size_t)(k0 + i) * N) : (seed ? p2c_dropout_scale(bits, k0 + i) : 1.f);
      v[i] = t * mq[i];
    }
    *reinterpret_cast<float4*>(out + k0) = make_float4(v[0], v[1], v[2], v[3]);
    if (A_out) {
      // the heads' input, relu(bn1(fc1)) * dropout mask, for their weight gradient (p2c_wgrad on the tensor cores)
      const float4 h = __ldg(reinterpret_cast<const float4*>(H + m * ldh + k0));
      float4 a;
      a.x = (scale ? fmaxf(fmaf(h.x, s_sc[k0], s_sh[k0]), 0.f) : h.x) * mq[0];
      a.y = (scale ? fmaxf(fmaf(h.y, s_sc[k0 + 1], s_sh[k0 + 1]), 0.f) : h.y) * mq[1];
      a.z = (scale ? fmaxf(fmaf(h.z, s_sc[k0 + 2], s_sh[k0 + 2]), 0.f) : h.z) * mq[2];
      a.w = (scale ? fmaxf(fmaf(h.w, s_sc[k0 + 3], s_sh[k0 + 3]), 0.f) : h.w) * mq[3];
      *reinterpret_cast<float4*>(A_out + m * lda + k0) = a;
    }
  }
}

// The same for C == 128 with the Philox mask (the training step's call): lane = four consecutive channels, so a warp
// reads / writes whole 512-byte rows (the kernel above is thread = row: 16-byte pieces of 32 different rows per
// instruction, half-used sectors - measured 281 us at 262,144 rows against ~62 us of HBM time).  The lane's 4 columns of
// W stay in registers; a row's Nout upstream gradients arrive by warp-uniform 16-byte loads; the keep-bits of 32 rows
// are drawn by the 32 lanes (one Philox call each) and handed round by shuffles.
template <int NP>
__global__ void __launch_bounds__(256, 2)
head_bwd_rows_kernel(const float* __restrict__ dOut, int64_t ldo, const int64_t* __restrict__ seed,
                     const float* __restrict__ W, int64_t M, int Nout, float* __restrict__ dA, int64_t ldda,
                     const float* __restrict__ H, int64_t ldh, const float* __restrict__ scale,
                     const float* __restrict__ shift, float* __restrict__ A_out, int64_t lda) {
  constexpr int C = 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k0 = lane * 4;
  float4 w[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j)
    w[j] = j < Nout ? __ldg(reinterpret_cast<const float4*>(W + (size_t)j * C + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (scale) {
    sc = __ldg(reinterpret_cast<const float4*>(scale + k0));
    sh = __ldg(reinterpret_cast<const float4*>(shift + k0));
  }
  const int word = lane >> 3, bit0 = (lane & 7) * 4;       // channel c: bit (c & 31) of word (c >> 5)
  __shared__ __align__(16) float s_g[8][32][NP];           // the warp's 32 upstream rows (columns >= Nout zeroed)
  for (int64_t m0 = ((int64_t)blockIdx.x * 8 + warp) * 32; m0 < M; m0 += (int64_t)gridDim.x * 256) {
    P2CPhilox4 bits;
    bits.v[0] = bits.v[1] = bits.v[2] = bits.v[3] = 0xffffffffu;
    if (seed && m0 + lane < M) bits = p2c_dropout_bits(seed, m0 + lane, 0);
    const int rows = (int)min((int64_t)32, M - m0);
    __syncwarp();
    if (lane < rows) {
      const float4* gp = reinterpret_cast<const float4*>(dOut + (m0 + lane) * ldo);
#pragma unroll
      for (int j4 = 0; j4 < NP / 4; ++j4) {
        float4 g = __ldg(gp + j4);
        // columns >= Nout are padding of the caller's buffer (uninitialised): they must not reach the FMAs as NaN * 0
        if (j4 * 4 + 1 >= Nout) g.y = 0.f;
        if (j4 * 4 + 2 >= Nout) g.z = 0.f;
        if (j4 * 4 + 3 >= Nout) g.w = 0.f;
        *reinterpret_cast<float4*>(&s_g[warp][lane][j4 * 4]) = g;
      }
    }
    __syncwarp();
    for (int r0 = 0; r0 < rows; r0 += 4) {
      float4 h[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {                         // four rows of H in flight before the arithmetic starts
        h[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (A_out && r0 + u < rows) h[u] = __ldg(reinterpret_cast<const float4*>(H + (m0 + r0 + u) * ldh + k0));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u;
        if (r >= rows) break;
        const int64_t m = m0 + r;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j4 = 0; j4 < NP / 4; ++j4) {
          const float4 g = *reinterpret_cast<const float4*>(&s_g[warp][r][j4 * 4]);   // broadcast read
          const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 ww = w[j4 * 4 + i];
            t.x = fmaf(gv[i], ww.x, t.x); t.y = fmaf(gv[i], ww.y, t.y);
            t.z = fmaf(gv[i], ww.z, t.z); t.w = fmaf(gv[i], ww.w, t.w);
          }
        }
        const uint32_t b0 = __shfl_sync(0xffffffffu, bits.v[0], r), b1 = __shfl_sync(0xffffffffu, bits.v[1], r);
        const uint32_t b2 = __shfl_sync(0xffffffffu, bits.v[2], r), b3 = __shfl_sync(0xffffffffu, bits.v[3], r);
        const uint32_t bw = (word == 0 ? b0 : word == 1 ? b1 : word == 2 ? b2 : b3) >> bit0;
        const float keep = seed ? 2.f : 1.f;
        const float4 mq = make_float4((bw & 1u) ? keep : 0.f, (bw & 2u) ? keep : 0.f, (bw & 4u) ? keep : 0.f,
                                      (bw & 8u) ? keep : 0.f);
        *reinterpret_cast<float4*>(dA + m * ldda + k0) = make_float4(t.x * mq.x, t.y * mq.y, t.z * mq.z, t.w * mq.w);
        if (A_out) {
          float4 a = h[u];
          if (scale) {
            a.x = fmaxf(fmaf(a.x, sc.x, sh.x), 0.f); a.y = fmaxf(fmaf(a.y, sc.y, sh.y), 0.f);
            a.z = fmaxf(fmaf(a.z, sc.z, sh.z), 0.f); a.w = fmaxf(fmaf(a.w, sc.w, sh.w), 0.f);
          }
          *reinterpret_cast<float4*>(A_out + m * lda + k0) = make_float4(a.x * mq.x, a.y * mq.y, a.z * mq.z, a.w * mq.w);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam defaults of train_...:189: betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad)
// over the flat parameter buffer: one launch for the whole model.
// ---------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            int64_t n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

}  // namespace

extern "C" int p2c_bn_bwd_reduce(const float* dA, int64_t ldda, const float* Y, int64_t ldy, const float* scale,
                                 const float* shift, int64_t M, int C, double* sums, void* stream) {
  if (!dA || !Y || !sums || M <= 0 || C <= 0 || ldda < C || ldy < C) return P2C_EINVAL;
  if ((scale == nullptr) != (shift == nullptr)) return P2C_EINVAL;
  if (C != 32 && C != 64 && (C % RED_CB) != 0) return P2C_EUNSUPPORTED;
  if ((ldda % 4) || (ldy % 4) ||
      ((reinterpret_cast<uintptr_t>(dA) | reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(scale) |
        reinterpret_cast<uintptr_t>(shift)) & 15))
    return P2C_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  P2C_CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
  const int cblocks = C < RED_CB ? 1 : C / RED_CB;
  // enough CTAs to fill the 148 SMs a few times over, at least 64 rows each
  int64_t rows_per_cta = M / ((int64_t)148 * 8 / cblocks + 1) + 1;
  if (rows_per_cta < 64) rows_per_cta = 64;
  if (rows_per_cta > 2048) rows_per_cta = 2048;
  dim3 grid(p2c_ceil_div(M, rows_per_cta), cblocks);
  bn_bwd_reduce_kernel<<<grid, RED_THREADS, 0, st>>>(dA, ldda, Y, ldy, scale, shift, M, C, (int)rows_per_cta, sums);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_pool_bwd_reduce(const float* dOut, int64_t ldd, const float* Ymax, const float* Ymin,
                                   const float* scale, const float* shift, int64_t G, int C, double* sums,
                                   void* stream) {
  if (!dOut || !Ymax || !Ymin || !scale || !shift || !sums || G <= 0 || C <= 0 || ldd < C) return P2C_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  P2C_CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
  dim3 grid(p2c_ceil_div(C, 128), p2c_ceil_div(G, 64));
  pool_bwd_reduce_kernel<<<grid, 128, 0, st>>>(dOut, ldd, Ymax, Ymin, scale, shift, G, C, sums);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_bn_bwd_coef(const double* sums, int64_t count, const float* gamma, const float* mean,
                               const float* invstd, int training, float* coef, float* dgamma, float* dbeta, int C,
                               void* stream) {
  if (!sums || !mean || !invstd || !coef || count <= 0 || C <= 0) return P2C_EINVAL;
  bn_bwd_coef_kernel<<<p2c_ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, (double)count, gamma, mean, invstd,
                                                                            training, coef, dgamma, dbeta, C);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_bn_bwd_apply(const float* dA, int64_t ldda, const float* Y, int64_t ldy, const float* scale,
                                const float* shift, const float* coef, int64_t M, int C, float* dY, int64_t lddy,
                                void* stream) {
  if (!dA || !Y || !scale || !shift || !coef || !dY || M <= 0 || C <= 0) return P2C_EINVAL;
  if (C % 4 || ldda % 4 || ldy % 4 || lddy % 4 ||
      ((reinterpret_cast<uintptr_t>(dA) | reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(dY) |
        reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift) | reinterpret_cast<uintptr_t>(coef)) & 15))
    return P2C_EALIGN;
  const int64_t total = M * (C / 4);
  const int blocks = (int)min((int64_t)148 * 16, (total + 255) / 256);
  bn_bwd_apply_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dA, ldda, Y, ldy, scale, shift, coef, M, C, dY, lddy);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_pool_bwd_apply(const float* dOut, int64_t ldd, const float* Ymax, const float* Ymin, const float* Y,
                                  int64_t ldy, const float* scale, const float* shift, const float* coef, int64_t G,
                                  int group, int C, float* dY, int64_t lddy, void* stream) {
  if (!dOut || !Ymax || !Ymin || !Y || !scale || !shift || !coef || !dY || G <= 0 || group <= 0 || C <= 0)
    return P2C_EINVAL;
  pool_bwd_apply_kernel<false><<<p2c_ceil_div(G * C, 256), 256, 0, (cudaStream_t)stream>>>(
      dOut, ldd, Ymax, Ymin, Y, ldy, scale, shift, coef, BnBwdCoefIn{}, G, group, C, dY, lddy);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

// p2c_bn_bwd_coef + p2c_bn_bwd_apply in one launch (see include/point2cyl.h)
extern "C" int p2c_bn_bwd_apply_fused(const float* dA, int64_t ldda, const float* Y, int64_t ldy, const float* scale,
                                      const float* shift, const double* sums, int64_t count, const float* gamma,
                                      const float* mean, const float* invstd, int training, float* dgamma,
                                      float* dbeta, int64_t M, int C, float* dY, int64_t lddy, void* stream) {
  if (!dA || !Y || !scale || !shift || !sums || !mean || !invstd || !dY || M <= 0 || C <= 0 || count <= 0)
    return P2C_EINVAL;
  if (C % 4 || ldda % 4 || ldy % 4 || lddy % 4 ||
      ((reinterpret_cast<uintptr_t>(dA) | reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(dY) |
        reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift)) & 15))
    return P2C_EALIGN;
  const int C4 = C / 4;
  if (256 % C4 != 0 && C4 % 256 != 0) return P2C_EUNSUPPORTED;     // a thread must stay on one channel group
  const int64_t total = M * C4;
  int blocks = (int)min((int64_t)148 * 16, (total + 255) / 256);
  const int per = C4 > 256 ? C4 / 256 : 1;                          // blocks per full set of channel groups
  blocks = (blocks + per - 1) / per * per;
  const BnBwdCoefIn ci{sums, (double)count, gamma, mean, invstd, training, dgamma, dbeta};
  if (C > 4096) return P2C_EUNSUPPORTED;
  bn_bwd_apply_fused_kernel<<<blocks, 256, 3 * C * sizeof(float), (cudaStream_t)stream>>>(dA, ldda, Y, ldy, scale, shift, ci, M, C,
                                                                                          dY, lddy);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_pool_bwd_apply_fused(const float* dOut, int64_t ldd, const float* Ymax, const float* Ymin,
                                        const float* Y, int64_t ldy, const float* scale, const float* shift,
                                        const double* sums, int64_t count, const float* gamma, const float* mean,
                                        const float* invstd, int training, float* dgamma, float* dbeta, int64_t G,
                                        int group, int C, float* dY, int64_t lddy, void* stream) {
  if (!dOut || !Ymax || !Ymin || !Y || !scale || !shift || !sums || !mean || !invstd || !dY || G <= 0 || group <= 0 ||
      C <= 0 || count <= 0)
    return P2C_EINVAL;
  const BnBwdCoefIn ci{sums, (double)count, gamma, mean, invstd, training, dgamma, dbeta};
  pool_bwd_apply_kernel<true><<<p2c_ceil_div(G * C, 256), 256, 0, (cudaStream_t)stream>>>(
      dOut, ldd, Ymax, Ymin, Y, ldy, scale, shift, nullptr, ci, G, group, C, dY, lddy);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

int p2c_wgrad_tc(const float* dY, int64_t lddy, const float* X, int64_t ldx, const float* in_scale,
                 const float* in_shift, int64_t M, int N, int K, float* dW, int64_t lddw, float* db, cudaStream_t st);

extern "C" int p2c_wgrad(const float* dY, int64_t lddy, const float* X, int64_t ldx, const float* in_scale,
                         const float* in_shift, const float* mask_cf, int mask_N, int64_t M, int N, int K, float* dW,
                         int64_t lddw, float* db, int precision, void* stream) {
  if (!dY || !X || !dW || M <= 0 || N <= 0 || K <= 0 || lddy < N || ldx < K || lddw < K) return P2C_EINVAL;
  if ((in_scale == nullptr) != (in_shift == nullptr)) return P2C_EINVAL;
  if (mask_cf && mask_N <= 0) return P2C_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == P2C_PREC_3XTF32 && !mask_cf) {
    const int rc = p2c_wgrad_tc(dY, lddy, X, ldx, in_scale, in_shift, M, N, K, dW, lddw, db, st);
    if (rc != P2C_EUNSUPPORTED) return rc;       // taken by the tensor-core kernel (or a real error)
  }
  const bool big = N >= 128 && K >= 128;
  const int TN = big ? 128 : 64, TK = big ? 128 : 64;
  const int gx = p2c_ceil_div(N, TN), gy = p2c_ceil_div(K, TK);
  // split the rows so that the grid fills the 148 SMs a few times over; at least 256 rows per CTA
  int64_t want = (int64_t)148 * 4 / ((int64_t)gx * gy);
  if (want < 1) want = 1;
  int64_t rows_per_cta = (M + want - 1) / want;
  if (rows_per_cta < 256) rows_per_cta = 256;
  rows_per_cta = (rows_per_cta + WG_RS - 1) / WG_RS * WG_RS;
  const int gz = p2c_ceil_div(M, rows_per_cta);
  dim3 grid(gx, gy, gz);
  if (big)
    wgrad_kernel<128, 128><<<grid, 256, 0, st>>>(dY, lddy, X, ldx, in_scale, in_shift, mask_cf, mask_N, M, N, K,
                                                 (int)rows_per_cta, dW, lddw, db);
  else
    wgrad_kernel<64, 64><<<grid, 256, 0, st>>>(dY, lddy, X, ldx, in_scale, in_shift, mask_cf, mask_N, M, N, K,
                                               (int)rows_per_cta, dW, lddw, db);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_sa_first_bwd(const float* dY, int64_t lddy, const float* xyz, const float* new_xyz,
                                const int64_t* idx, int B, int N, int S, int nsample, int C, float* dQf, int64_t ldq,
                                float* dW, int64_t lddw, float* dbias, void* stream) {
  if (!dY || !xyz || !new_xyz || !idx || !dW || B <= 0 || N <= 0 || S <= 0 || nsample <= 0 || lddw < 3 || lddy < C)
    return P2C_EINVAL;
  if (C != 64 && C != 128) return P2C_EUNSUPPORTED;
  if ((lddy % 4) != 0 || (reinterpret_cast<uintptr_t>(dY) & 15) != 0) return P2C_EALIGN;
  const int64_t rows = (int64_t)B * S * nsample;
  if (rows >= ((int64_t)1 << 31) || (int64_t)B * N >= ((int64_t)1 << 32)) return P2C_EUNSUPPORTED;
  const int blocks = (int)min((int64_t)148 * 8, (rows + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 128)
    sa_first_bwd_kernel<4><<<blocks, 256, 0, st>>>(dY, lddy, xyz, new_xyz, idx, N, S, nsample, rows, dQf, ldq, dW, lddw,
                                                   dbias);
  else
    sa_first_bwd_kernel<2><<<blocks, 256, 0, st>>>(dY, lddy, xyz, new_xyz, idx, N, S, nsample, rows, dQf, ldq, dW, lddw,
                                                   dbias);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_group_bwd(const float* dRows, int64_t ldr, const int64_t* idx, int B, int N, int S, int nsample,
                             int D, float* dfeats, int64_t ldf, void* stream) {
  if (!dRows || !idx || !dfeats || B <= 0 || N <= 0 || S <= 0 || nsample <= 0 || D <= 0 || ldr < 3 + D || ldf < D)
    return P2C_EINVAL;
  const int64_t rows = (int64_t)B * S * nsample;
  group_bwd_kernel<<<p2c_ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(dRows, ldr, idx, N, S, nsample, D, rows,
                                                                           dfeats, ldf);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_three_nn_interp_bwd(const float* dInterp, int64_t ldi, const int64_t* idx, const float* w, int B,
                                       int N, int S, int D, float* dfeats2, int64_t ldf, void* stream) {
  if (!dInterp || !dfeats2 || B <= 0 || N <= 0 || S <= 0 || D <= 0 || ldi < D || ldf < D) return P2C_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (S == 1) {
    broadcast_bwd_kernel<<<dim3(p2c_ceil_div(D, 32), B), 256, 0, st>>>(dInterp, ldi, N, D, dfeats2, ldf);
  } else {
    if (!idx || !w) return P2C_EINVAL;
    P2C_CUDA_TRY(cudaMemset2DAsync(dfeats2, sizeof(float) * ldf, 0, sizeof(float) * D, (size_t)B * S, st));
    const int64_t rows = (int64_t)B * N;
    interp_bwd_kernel<<<p2c_ceil_div(rows, 8), 256, 0, st>>>(dInterp, ldi, idx, w, N, S, D, rows, dfeats2, ldf);
  }
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_head_bwd(const float* dOut, int64_t ldo, const float* mask_cf, const int64_t* dropout_seed,
                            const float* W, int B, int N, int C,
                            int Nout, float* dA, int64_t ldda, const float* H, int64_t ldh, const float* scale,
                            const float* shift, float* A_out, int64_t lda, void* stream) {
  if (!dOut || !W || !dA || B <= 0 || N <= 0 || C <= 0 || Nout <= 0 || ldo < Nout || ldda < C) return P2C_EINVAL;
  if ((scale == nullptr) != (shift == nullptr)) return P2C_EINVAL;
  if (A_out && (!H || ldh < C || lda < C)) return P2C_EINVAL;
  if (C % 4 || ldda % 4 || (reinterpret_cast<uintptr_t>(dA) & 15)) return P2C_EALIGN;
  if (A_out && (ldh % 4 || lda % 4 || ((reinterpret_cast<uintptr_t>(H) | reinterpret_cast<uintptr_t>(A_out)) & 15)))
    return P2C_EALIGN;
  if (C > 256 || Nout > 36) return P2C_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t M = (int64_t)B * N;
  // rows-as-warps kernel: 128 channels, Philox (or no) mask, upstream rows readable as 16-byte pieces up to the padded width
  if (C == 128 && !mask_cf && (ldo % 4) == 0 && (reinterpret_cast<uintptr_t>(dOut) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(W) & 15) == 0 && (!scale || ((reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift)) & 15) == 0)) {
    const int np = (Nout + 3) & ~3;
    if (np <= ldo && np <= 20) {
      const int blocks = (int)min((int64_t)148 * 16, (int64_t)p2c_ceil_div(M, 256));
#define P2C_HBR(NPV) head_bwd_rows_kernel<NPV><<<blocks, 256, 0, st>>>(dOut, ldo, dropout_seed, W, M, Nout, dA, ldda, H, ldh, scale, shift, A_out, lda)
      if (np <= 4) P2C_HBR(4); else if (np <= 8) P2C_HBR(8); else if (np <= 12) P2C_HBR(12);
      else if (np <= 16) P2C_HBR(16); else P2C_HBR(20);
#undef P2C_HBR
      P2C_RETURN_IF_CUDA_ERROR();
      return 0;
    }
  }
#define P2C_HB(NPV)                                                                                              \
  do {                                                                                                           \
    const size_t smem = ((size_t)C * NPV + 2 * C) * sizeof(float);                                               \
    head_bwd_kernel<NPV><<<p2c_ceil_div(M, 256), 256, smem, st>>>(dOut, ldo, mask_cf, dropout_seed, W, M, N, C, Nout,  \
                                                                  dA, ldda, H, ldh, scale, shift, A_out, lda);     \
  } while (0)
  if (Nout <= 4) P2C_HB(4); else if (Nout <= 8) P2C_HB(8); else if (Nout <= 12) P2C_HB(12);
  else if (Nout <= 20) P2C_HB(20); else if (Nout <= 28) P2C_HB(28); else P2C_HB(36);
#undef P2C_HB
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                             void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || n <= 0 || step <= 0) return P2C_EINVAL;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  const int blocks = (int)min((int64_t)148 * 8, (n + 255) / 256);
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                        weight_decay, bc1, bc2_sqrt, grad_scale);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
