// Per-point MLP layer, fp32 SIMT path (P2C_PREC_FP32), plus the BatchNorm bookkeeping kernels.
//
// Y = f(X) W^T + bias with the previous layer's BN+ReLU (and dropout mask) applied while the A tile
// is loaded, and this layer's per-channel sum / sum-of-squares and the nsample max/min pool reduced
// in the epilogue, so an activation crosses HBM once per layer (raw, pre-BN) instead of three
// times (conv out, BN out, ReLU out) as in the eager reference (models/pointnet_util.py:200-205).
// Classic 128 x BN x 16 register-tiled GEMM, 256 threads, 8 x BN/16 outputs per thread, double
// buffered through shared memory.  The tensor-core paths (tcgen05) live in linear_tc.cu; this
// one is the full-fp32 fallback for shapes they do not take and the in-library cross-check.
#include <cstdlib>

#include "bn_fold.cuh"

#include <math_constants.h>

namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int PADM = 4;

struct LinearArgs {
  const float* X; int64_t ldx;
  const float* W; const float* bias;
  const float* in_scale; const float* in_shift;
  const float* in_mask; int64_t ldmask;
  float* Y; int64_t ldy;
  int M, N, K;
  double* stats;
  int pool_group;
  float* Ymax; float* Ymin;
};

template <int BN>
struct Smem {
  union {
    struct {
      float A[2][BK][BM + PADM];
      float B[2][BK][BN + PADM];
    } t;
    float red[2][2][16][BN];  // [max|sum , min|sumsq][row chunk][ty][col]
  };
};

template <int BN, bool VEC>
__global__ void __launch_bounds__(256)
linear_simt_kernel(const LinearArgs a) {
  constexpr int TN = BN / 16;              // columns per thread
  constexpr int CW = TN >= 4 ? 4 : TN;     // column chunk width
  constexpr int NCH = TN / CW;             // column chunks per thread
  constexpr int WE = BN / 16;              // W elements per thread per k-tile
  constexpr int WKG = 256 / BN;            // k-groups in the W tile load
  __shared__ __align__(16) Smem<BN> sm;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int M = a.M, N = a.N, K = a.K;

  // ---- global -> register tile loaders --------------------------------------------------
  const int arow = tid & 127, akh = tid >> 7;
  const int wn = tid % BN, wkq = tid / BN;
  float ra[8], rw[WE];

  auto load_tiles = [&](int k0) {
    const int row = m0 + arow;
    const int kb = k0 + akh * 8;
    const bool rok = row < M;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = kb + h * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (VEC) {
        if (rok && kk < K) {
          float4 t = __ldg(reinterpret_cast<const float4*>(a.X + (size_t)row * a.ldx + kk));
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          if (a.in_scale) {
            float4 s = __ldg(reinterpret_cast<const float4*>(a.in_scale + kk));
            float4 f = __ldg(reinterpret_cast<const float4*>(a.in_shift + kk));
            v[0] = fmaxf(fmaf(v[0], s.x, f.x), 0.f); v[1] = fmaxf(fmaf(v[1], s.y, f.y), 0.f);
            v[2] = fmaxf(fmaf(v[2], s.z, f.z), 0.f); v[3] = fmaxf(fmaf(v[3], s.w, f.w), 0.f);
          }
          if (a.in_mask) {
            float4 mk = __ldg(reinterpret_cast<const float4*>(a.in_mask + (size_t)row * a.ldmask + kk));
            v[0] *= mk.x; v[1] *= mk.y; v[2] *= mk.z; v[3] *= mk.w;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = kk + i;
          if (rok && k < K) {
            float x = __ldg(a.X + (size_t)row * a.ldx + k);
            if (a.in_scale) x = fmaxf(fmaf(x, __ldg(a.in_scale + k), __ldg(a.in_shift + k)), 0.f);
            if (a.in_mask) x *= __ldg(a.in_mask + (size_t)row * a.ldmask + k);
            v[i] = x;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) ra[h * 4 + i] = v[i];
    }
    const int n = n0 + wn;
    const int wk = k0 + wkq * WE;
#pragma unroll
    for (int i = 0; i < WE; ++i) rw[i] = 0.f;
    if (n < N) {
      const float* wp = a.W + (size_t)n * K + wk;
      if (VEC && WE >= 4) {
#pragma unroll
        for (int i = 0; i < WE; i += 4) {
          if (wk + i < K) {
            float4 t = __ldg(reinterpret_cast<const float4*>(wp + i));
            rw[i] = t.x; rw[i + 1] = t.y; rw[i + 2] = t.z; rw[i + 3] = t.w;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < WE; ++i)
          if (wk + i < K) rw[i] = __ldg(wp + i);
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sm.t.A[buf][akh * 8 + i][arow] = ra[i];
#pragma unroll
    for (int i = 0; i < WE; ++i) sm.t.B[buf][wkq * WE + i][wn] = rw[i];
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int ktiles = (K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ktiles) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[8], bv[TN];
      float4 t0 = *reinterpret_cast<const float4*>(&sm.t.A[buf][k][ty * 4]);
      float4 t1 = *reinterpret_cast<const float4*>(&sm.t.A[buf][k][64 + ty * 4]);
      av[0] = t0.x; av[1] = t0.y; av[2] = t0.z; av[3] = t0.w;
      av[4] = t1.x; av[5] = t1.y; av[6] = t1.z; av[7] = t1.w;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (CW == 4) {
          float4 u = *reinterpret_cast<const float4*>(&sm.t.B[buf][k][c * (BN / NCH) + tx * 4]);
          bv[c * 4 + 0] = u.x; bv[c * 4 + 1] = u.y; bv[c * 4 + 2] = u.z; bv[c * 4 + 3] = u.w;
        } else {
#pragma unroll
          for (int j = 0; j < CW; ++j) bv[c * CW + j] = sm.t.B[buf][k][tx * CW + j];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < ktiles) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue ---------------------------------------------------------------------------
  auto col_of = [&](int j) { return n0 + (j / CW) * (BN / NCH) + tx * CW + (j % CW); };
  auto lcol_of = [&](int j) { return (j / CW) * (BN / NCH) + tx * CW + (j % CW); };
  auto row_of = [&](int i) { return m0 + (i >> 2) * 64 + ty * 4 + (i & 3); };

#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int c = col_of(j);
    const float bj = (a.bias && c < N) ? __ldg(a.bias + c) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][j] += bj;
  }

  if (a.Y) {
    const bool vecy = (CW == 4) && (a.ldy % 4 == 0) && (N % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(a.Y) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = row_of(i);
      if (r >= M) continue;
      float* yr = a.Y + (size_t)r * a.ldy;
      if (vecy) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int col = col_of(c * 4);
          if (col < N)
            *reinterpret_cast<float4*>(yr + col) =
                make_float4(acc[i][c * 4], acc[i][c * 4 + 1], acc[i][c * 4 + 2], acc[i][c * 4 + 3]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const int col = col_of(j);
          if (col < N) yr[col] = acc[i][j];
        }
      }
    }
  }

  if (a.stats || a.pool_group) __syncthreads();  // everyone is done with the A/B tiles

  if (a.stats) {
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      float s = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float y = row_of(i) < M ? acc[i][j] : 0.f;
        s += y;
        s2 = fmaf(y, y, s2);
      }
      sm.red[0][0][ty][lcol_of(j)] = s;
      sm.red[1][0][ty][lcol_of(j)] = s2;
    }
    __syncthreads();
    if (tid < BN && n0 + tid < N) {
      double s = 0.0, s2 = 0.0;
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        s += (double)sm.red[0][0][t][tid];
        s2 += (double)sm.red[1][0][t][tid];
      }
      atomicAdd(a.stats + n0 + tid, s);
      atomicAdd(a.stats + N + n0 + tid, s2);
    }
    if (a.pool_group) __syncthreads();
  }

  if (a.pool_group) {
    const int G = a.pool_group;  // divides 128, multiple of 4 (checked by the host)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        float mx = -CUDART_INF_F, mn = CUDART_INF_F;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (row_of(h * 4 + i) < M) {
            mx = fmaxf(mx, acc[h * 4 + i][j]);
            mn = fminf(mn, acc[h * 4 + i][j]);
          }
        }
        sm.red[0][h][ty][lcol_of(j)] = mx;
        sm.red[1][h][ty][lcol_of(j)] = mn;
      }
    }
    __syncthreads();
    const int groups = BM / G;
    for (int e = tid; e < groups * BN; e += 256) {
      const int g = e / BN, c = e % BN;
      const int grow = m0 + g * G;
      if (grow >= M || n0 + c >= N) continue;
      float mx = -CUDART_INF_F, mn = CUDART_INF_F;
      // thread-row-chunks (h, t) cover tile rows h*64 + t*4 .. +3
      for (int r = g * G; r < (g + 1) * G; r += 4) {
        const int h = r >> 6, t = (r & 63) >> 2;
        mx = fmaxf(mx, sm.red[0][h][t][c]);
        mn = fminf(mn, sm.red[1][h][t][c]);
      }
      const size_t o = (size_t)(grow / G) * N + n0 + c;
      a.Ymax[o] = mx;
      a.Ymin[o] = mn;
    }
  }
}

template <int BN>
int launch_linear(const LinearArgs& a, cudaStream_t st) {
  dim3 grid(p2c_ceil_div(a.M, BM), p2c_ceil_div(a.N, BN));
  const bool vec = (a.K % 4 == 0) && (a.ldx % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(a.X) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(a.W) & 15) == 0) &&
                   (!a.in_scale || (((reinterpret_cast<uintptr_t>(a.in_scale) |
                                      reinterpret_cast<uintptr_t>(a.in_shift)) & 15) == 0)) &&
                   (!a.in_mask || ((a.ldmask % 4 == 0) &&
                                   (reinterpret_cast<uintptr_t>(a.in_mask) & 15) == 0));
  if (vec)
    linear_simt_kernel<BN, true><<<grid, 256, 0, st>>>(a);
  else
    linear_simt_kernel<BN, false><<<grid, 256, 0, st>>>(a);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

// ---- BatchNorm bookkeeping ----------------------------------------------------------------------

__global__ void bn_finalize_kernel(const double* __restrict__ stats, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float eps, float momentum, int training, float* running_mean,
                                   float* running_var, float* __restrict__ scale,
                                   float* __restrict__ shift, float* save_mean, float* save_invstd,
                                   int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mean, var;
  if (training) {
    mean = stats[c] / count;
    var = stats[C + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    if (running_mean) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * mean);
      running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * unbiased);
    }
  } else {
    mean = (double)running_mean[c];
    var = (double)running_var[c];
  }
  const double invstd = 1.0 / sqrt(var + (double)eps);
  const double g = gamma ? (double)gamma[c] : 1.0;
  const double bta = beta ? (double)beta[c] : 0.0;
  const float sc = (float)(g * invstd);
  scale[c] = sc;
  shift[c] = (float)(bta - mean * g * invstd);
  if (save_mean) save_mean[c] = (float)mean;
  if (save_invstd) save_invstd[c] = (float)invstd;
}

// The pending BatchNorm folded by the consumer: a CTA owns a block of 32 CHANNELS (blockIdx.x) and a range of rows
// (blockIdx.y), so it folds only its 32 channels - one float64 division and square root per lane of its first warp -
// instead of all C (with 2368 CTAs folding up to 1024 channels each these two kernels took 12-15 us for a few MB).
// Thread = (channel lane, row lane): a warp covers the 32 channels of one row, one 128-byte line.
__device__ __forceinline__ void fold_block(const BnFoldDev& bn, int C, float& sc, float& sh) {
  __shared__ float s_f[2][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  if (threadIdx.x < 32) {
    float a = 0.f, b = 0.f;
    if (c < C) p2c_bn_fold_channel(bn, c, blockIdx.y == 0, a, b);
    s_f[0][threadIdx.x] = a;
    s_f[1][threadIdx.x] = b;
  }
  __syncthreads();
  sc = s_f[0][threadIdx.x & 31];
  sh = s_f[1][threadIdx.x & 31];
}

__global__ void __launch_bounds__(256)
bn_relu_apply_fold_kernel(const float* __restrict__ Y, int64_t ldy, const BnFoldDev bn, float* __restrict__ out,
                          int64_t ldo, int64_t M, int C, int64_t rows_per_cta) {
  float sc, sh;
  fold_block(bn, C, sc, sh);
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  if (c >= C) return;
  const int64_t m1 = min(M, ((int64_t)blockIdx.y + 1) * rows_per_cta);
  for (int64_t m = (int64_t)blockIdx.y * rows_per_cta + (threadIdx.x >> 5); m < m1; m += 8)
    out[m * ldo + c] = fmaxf(fmaf(__ldg(Y + m * ldy + c), sc, sh), 0.f);
}

__global__ void __launch_bounds__(256)
pool_bn_relu_fold_kernel(const float* __restrict__ Ymax, const float* __restrict__ Ymin, const BnFoldDev bn,
                         float* __restrict__ out, int64_t ldo, int64_t G, int C, int64_t rows_per_cta) {
  float sc, sh;
  fold_block(bn, C, sc, sh);
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  if (c >= C) return;
  const float* src = sc >= 0.f ? Ymax : Ymin;
  const int64_t g1 = min(G, ((int64_t)blockIdx.y + 1) * rows_per_cta);
  for (int64_t g = (int64_t)blockIdx.y * rows_per_cta + (threadIdx.x >> 5); g < g1; g += 8)
    out[g * ldo + c] = fmaxf(fmaf(__ldg(src + g * C + c), sc, sh), 0.f);
}

// grid of the two kernels above: (channel blocks, row ranges), about four CTAs per SM in total
static inline dim3 fold_grid(int64_t rows, int C, int64_t* rows_per_cta) {
  const int cb = (C + 31) / 32;
  int64_t ry = (592 + cb - 1) / cb;
  if (ry > (rows + 7) / 8) ry = (rows + 7) / 8;
  if (ry < 1) ry = 1;
  int64_t rpc = (rows + ry - 1) / ry;
  rpc = (rpc + 7) / 8 * 8;
  *rows_per_cta = rpc;
  return dim3((unsigned)cb, (unsigned)((rows + rpc - 1) / rpc));
}

__global__ void __launch_bounds__(256)
bn_relu_apply_kernel(const float* __restrict__ Y, int64_t ldy, const float* __restrict__ scale,
                     const float* __restrict__ shift, float* __restrict__ out, int64_t ldo,
                     int64_t M, int C) {
  const int64_t total = M * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = e / C;
    const int c = (int)(e - m * C);
    out[m * ldo + c] = fmaxf(fmaf(__ldg(Y + m * ldy + c), __ldg(scale + c), __ldg(shift + c)), 0.f);
  }
}

__global__ void __launch_bounds__(256)
pool_bn_relu_kernel(const float* __restrict__ Ymax, const float* __restrict__ Ymin,
                    const float* __restrict__ scale, const float* __restrict__ shift,
                    float* __restrict__ out, int64_t ldo, int64_t G, int C) {
  const int64_t total = G * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = e / C;
    const int c = (int)(e - g * C);
    const float s = __ldg(scale + c);
    const float v = s >= 0.f ? __ldg(Ymax + e) : __ldg(Ymin + e);
    out[g * ldo + c] = fmaxf(fmaf(v, s, __ldg(shift + c)), 0.f);
  }
}

// A layer over a handful of rows (one per cloud): four output channels per CTA with their weight rows in shared
// memory, warp = row (blockIdx.y picks the group of eight rows), lanes stride over K with eight loads in flight.
// The tiled kernels above would run it on one or two CTAs.
__global__ void __launch_bounds__(256)
linear_small_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ W, int64_t ldw,
                    const float* __restrict__ bias, float* __restrict__ Y, int64_t ldy, int M, int N, int K) {
  extern __shared__ float s_w4[];              // [4][K]
  const int n0 = blockIdx.x * 4;
  for (int e = threadIdx.x; e < 4 * K; e += 256) {
    const int j = e / K, k = e - j * K;
    s_w4[e] = (n0 + j < N) ? __ldg(W + (size_t)(n0 + j) * ldw + k) : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.y * 8 + warp;
  if (m >= M) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const float* xr = X + (size_t)m * ldx;
  for (int k0 = 0; k0 < K; k0 += 256) {
    float x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = k0 + u * 32 + lane;
      x[u] = k < K ? __ldg(xr + k) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = k0 + u * 32 + lane;
      if (k < K) {
        a0 = fmaf(x[u], s_w4[k], a0); a1 = fmaf(x[u], s_w4[K + k], a1);
        a2 = fmaf(x[u], s_w4[2 * K + k], a2); a3 = fmaf(x[u], s_w4[3 * K + k], a3);
      }
    }
  }
  a0 = p2c_warp_sum(a0); a1 = p2c_warp_sum(a1); a2 = p2c_warp_sum(a2); a3 = p2c_warp_sum(a3);
  if (lane < 4 && n0 + lane < N) {
    const float v = lane == 0 ? a0 : lane == 1 ? a1 : lane == 2 ? a2 : a3;
    Y[(size_t)m * ldy + n0 + lane] = v + (bias ? __ldg(bias + n0 + lane) : 0.f);
  }
}

}  // namespace

extern "C" int p2c_linear_small(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* bias, float* Y,
                                int64_t ldy, int M, int N, int K, void* stream) {
  if (!X || !W || !Y || M <= 0 || N <= 0 || K <= 0 || ldx < K || ldw < K || ldy < N) return P2C_EINVAL;
  if (M > 256 || K > 3072) return P2C_EUNSUPPORTED;
  const dim3 grid(p2c_ceil_div(N, 4), p2c_ceil_div(M, 8));
  linear_small_kernel<<<grid, 256, 4 * K * sizeof(float), (cudaStream_t)stream>>>(X, ldx, W, ldw, bias, Y, ldy, M, N, K);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

// tensor-core paths, linear_tc.cu
int p2c_linear_tc(const float* X, int64_t ldx, const float* W, const float* bias,
                  const float* in_scale, const float* in_shift, const float* in_mask,
                  int64_t ldmask, float* Y, int64_t ldy, int M, int N, int K, double* stats,
                  int pool_group, float* Ymax, float* Ymin, int precision, const p2c_bn_fold* in_bn, cudaStream_t st,
                  const int64_t* drop_seed = nullptr, const P2cXyzFirst* xyz_first = nullptr);
int p2c_linear_tc_ss_plan(int64_t ldx, int x_aligned16, int K, int has_mask, int pool_group, int precision);
int p2c_linear_tc_ss(const float* X, int64_t ldx, const float* w_split, int64_t ldws, const float* bias,
                     const float* in_scale, const float* in_shift, float* Y, int64_t ldy, int M, int N, int K,
                     double* stats, int pool_group, float* Ymax, float* Ymin, int bf16, const p2c_bn_fold* in_bn,
                     cudaStream_t st);

// the stand-alone finalisation of a pending BatchNorm (kernels that do not fold it themselves)
static int resolve_bn_fold(const p2c_bn_fold* f, cudaStream_t st) {
  bn_finalize_kernel<<<p2c_ceil_div(f->C, 128), 128, 0, st>>>(f->stats, (double)f->count, f->gamma, f->beta, f->eps,
                                                            f->momentum, f->stats != nullptr, f->running_mean,
                                                            f->running_var, f->scale_out, f->shift_out, f->mean_out,
                                                            f->invstd_out, f->C);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_linear(const float* X, int64_t ldx, const float* W, const float* bias,
                          const float* in_scale, const float* in_shift, const float* in_mask,
                          int64_t ldmask, float* Y, int64_t ldy, int M, int N, int K, double* stats,
                          int pool_group, float* Ymax, float* Ymin, int precision, const float* w_split,
                          int64_t ldws, const p2c_bn_fold* in_bn, void* stream) {
  if (!X || !W || M <= 0 || N <= 0 || K <= 0 || ldx < K) return P2C_EINVAL;
  if ((in_scale == nullptr) != (in_shift == nullptr)) return P2C_EINVAL;
  if (int e = p2c_bn_fold_check(in_bn, K)) return e;
  if (!Y && !pool_group && !stats) return P2C_EINVAL;
  if (Y && ldy < N) return P2C_EINVAL;
  if (pool_group) {
    if (!Ymax || !Ymin) return P2C_EINVAL;
    if (pool_group < 4 || BM % pool_group != 0 || M % pool_group != 0) return P2C_EUNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (precision != P2C_PREC_FP32) {
    static const bool force_ss = getenv("P2C_TC_FORCE_SS") != nullptr;   // tools only: streamed-weight kernel first
    int rc = P2C_EUNSUPPORTED;
    if (!(force_ss && w_split) && precision != P2C_PREC_BF16)
      rc = p2c_linear_tc(X, ldx, W, bias, in_scale, in_shift, in_mask, ldmask, Y, ldy, M, N, K,
                         stats, pool_group, Ymax, Ymin, precision, in_bn, st);
    if (rc != P2C_EUNSUPPORTED) return rc;
    if (w_split && p2c_linear_tc_ss_plan(ldx, (reinterpret_cast<uintptr_t>(X) & 15) == 0, K, in_mask != nullptr,
                                         pool_group, precision)) {
      rc = p2c_linear_tc_ss(X, ldx, w_split, ldws, bias, in_scale, in_shift, Y, ldy, M, N, K, stats, pool_group,
                            Ymax, Ymin, precision == P2C_PREC_BF16 ? 1 : 0, in_bn, st);
      if (rc != P2C_EUNSUPPORTED) return rc;
    }
    // shapes the tensor-core kernels do not take fall through to the fp32 SIMT kernel (still CUDA)
  }
  if (in_bn) {   // the SIMT kernel reads scale / shift from global memory: finalise first
    if (int e = resolve_bn_fold(in_bn, st)) return e;
    in_scale = in_bn->scale_out;
    in_shift = in_bn->shift_out;
  }
  LinearArgs a{X, ldx, W, bias, in_scale, in_shift, in_mask, ldmask, Y, ldy, M, N, K,
               stats, pool_group, Ymax, Ymin};
  if (N > 64) return launch_linear<128>(a, st);
  if (N > 32) return launch_linear<64>(a, st);
  return launch_linear<32>(a, st);
}

extern "C" int p2c_bn_finalize(const double* stats, int64_t count, const float* gamma,
                               const float* beta, float eps, float momentum, int training,
                               float* running_mean, float* running_var, float* scale, float* shift,
                               float* save_mean, float* save_invstd, int C, void* stream) {
  if (!scale || !shift || C <= 0) return P2C_EINVAL;
  if (training && (!stats || count <= 0)) return P2C_EINVAL;
  if (!training && (!running_mean || !running_var)) return P2C_EINVAL;
  bn_finalize_kernel<<<p2c_ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(
      stats, (double)count, gamma, beta, eps, momentum, training, running_mean, running_var, scale,
      shift, save_mean, save_invstd, C);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_bn_relu_apply(const float* Y, int64_t ldy, const float* scale, const float* shift,
                                 float* out, int64_t ldo, int64_t M, int C, const p2c_bn_fold* bn, void* stream) {
  if (!Y || !out || M <= 0 || C <= 0 || (!bn && (!scale || !shift))) return P2C_EINVAL;
  if (int e = p2c_bn_fold_check(bn, C)) return e;
  const int64_t total = M * C;
  const int blocks = (int)min((int64_t)148 * 16, (total + 255) / 256);
  if (bn) {
    int64_t rpc;
    const dim3 grid = fold_grid(M, C, &rpc);
    bn_relu_apply_fold_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Y, ldy, p2c_bn_fold_dev(bn), out, ldo, M, C, rpc);
    P2C_RETURN_IF_CUDA_ERROR();
    return 0;
  }
  bn_relu_apply_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(Y, ldy, scale, shift, out, ldo, M, C);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_pool_bn_relu(const float* Ymax, const float* Ymin, const float* scale,
                                const float* shift, float* out, int64_t ldo, int64_t G, int C,
                                const p2c_bn_fold* bn, void* stream) {
  if (!Ymax || !Ymin || !out || G <= 0 || C <= 0 || (!bn && (!scale || !shift))) return P2C_EINVAL;
  if (int e = p2c_bn_fold_check(bn, C)) return e;
  const int64_t total = G * C;
  const int blocks = (int)min((int64_t)148 * 16, (total + 255) / 256);
  if (bn) {
    int64_t rpc;
    const dim3 grid = fold_grid(G, C, &rpc);
    pool_bn_relu_fold_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Ymax, Ymin, p2c_bn_fold_dev(bn), out, ldo, G, C, rpc);
    P2C_RETURN_IF_CUDA_ERROR();
    return 0;
  }
  pool_bn_relu_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(Ymax, Ymin, scale, shift, out, ldo, G, C);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
