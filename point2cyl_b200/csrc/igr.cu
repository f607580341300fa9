// Small kernels around the implicit sketch network (IGR/network.py, train_Point2Cyl.py:598-672); the 512-wide hidden
// layers themselves run on the tensor cores (p2c_linear_act, linear_tc_ss.cu).
//   p2c_igr_add_latent   add_latent (IGR/network.py:200-206): rows [latent code | 2-D point], plus the copy of that row
//                        that the skip layer concatenates (network.py:80-81), already scaled by 1/sqrt(2)
//   p2c_igr_scale_cols   out = A * w[col]: seed of the reverse sweep, a_L-1 = softplus'(z_L-1) * W_L
//   p2c_igr_rowdots      out[m, j] (+)= scale * <A[m, :], V[j, :]> + bias[j]: the 512 -> 1 output layer and the two
//                        trailing columns of the input gradient (gradient() keeps only the 2-D point's columns, :17)
//   p2c_igr_loss_terms   per sketch instance: mean |f|, mean min(|g - n|, |g + n|), mean (|g_off| - 1)^2
//                        (manifold / SALD-normal / eikonal terms, train_Point2Cyl.py:627-647)
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
igr_add_latent_kernel(const float* __restrict__ latent, const float* __restrict__ pts, int64_t R, int S, int E,
                      float* __restrict__ X0, int64_t ldx, float* __restrict__ P, int64_t ldp, int colp, float pscale) {
  const int W = E + 2;
  const int64_t total = R * ldx;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / ldx;
    const int c = (int)(e - r * ldx);
    float v = 0.f;
    if (c < E) v = __ldg(latent + (r / S) * E + c);
    else if (c < W) v = __ldg(pts + r * 2 + (c - E));
    X0[e] = v;
    if (P && c < W) P[r * ldp + colp + c] = v * pscale;
  }
}

__global__ void __launch_bounds__(256)
igr_scale_cols_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ w, int64_t M, int C,
                      float* __restrict__ out, int64_t ldo) {
  const int C4 = C >> 2;
  const int64_t total = M * C4;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = e / C4;
    const int c = (int)(e - m * C4) * 4;
    const float4 a = __ldg(reinterpret_cast<const float4*>(A + m * lda + c));
    const float4 v = __ldg(reinterpret_cast<const float4*>(w + c));
    *reinterpret_cast<float4*>(out + m * ldo + c) = make_float4(a.x * v.x, a.y * v.y, a.z * v.z, a.w * v.w);
  }
}

// warp per row; V (NV, C) in shared memory
template <int NV>
__global__ void __launch_bounds__(256)
igr_rowdots_kernel(const float* __restrict__ A, int64_t lda, int64_t M, int C, const float* __restrict__ V,
                   int64_t vsj, int64_t vsc, const float* __restrict__ bias, float scale, float* __restrict__ out,
                   int64_t ldo, int accumulate) {
  extern __shared__ __align__(16) float s_v[];
  for (int e = threadIdx.x; e < NV * C; e += blockDim.x) s_v[e] = __ldg(V + (e / C) * vsj + (e % C) * vsc);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t m = warp; m < M; m += nwarps) {
    float acc[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] = 0.f;
    const float* a = A + m * lda;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(a + c));
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(s_v + j * C + c);
        acc[j] = fmaf(x.x, v.x, fmaf(x.y, v.y, fmaf(x.z, v.z, fmaf(x.w, v.w, acc[j]))));
      }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float t = p2c_warp_sum(acc[j]);
      if (lane == 0) {
        float r = t * scale + (bias ? __ldg(bias + j) : 0.f);
        if (accumulate) r += out[m * ldo + j];
        out[m * ldo + j] = r;
      }
    }
  }
}

__device__ __forceinline__ float block_sum_256(float v, float* s_red) {
  v = p2c_warp_sum(v);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < 8 ? s_red[threadIdx.x] : 0.f;
  t = p2c_warp_sum(t);
  __syncthreads();
  return t;                       // valid in warp 0
}

// one CTA per sketch instance
__global__ void __launch_bounds__(256)
igr_loss_terms_kernel(const float* __restrict__ f_on, const float* __restrict__ g_on, const float* __restrict__ nrm,
                      const float* __restrict__ g_off, int S, int So, float* __restrict__ out) {
  __shared__ float s_red[8];
  const int64_t i = blockIdx.x;
  float a = 0.f, b = 0.f, c = 0.f;
  for (int p = threadIdx.x; p < S; p += 256) {
    const int64_t r = i * S + p;
    a += fabsf(__ldg(f_on + r));
    const float gx = __ldg(g_on + 2 * r), gy = __ldg(g_on + 2 * r + 1);
    const float nx = __ldg(nrm + 2 * r), ny = __ldg(nrm + 2 * r + 1);
    const float sub = sqrtf((gx - nx) * (gx - nx) + (gy - ny) * (gy - ny));
    const float add = sqrtf((gx + nx) * (gx + nx) + (gy + ny) * (gy + ny));
    b += fminf(sub, add);
  }
  for (int p = threadIdx.x; p < So; p += 256) {
    const int64_t r = i * So + p;
    const float gx = __ldg(g_off + 2 * r), gy = __ldg(g_off + 2 * r + 1);
    const float d = sqrtf(gx * gx + gy * gy) - 1.f;
    c += d * d;
  }
  a = block_sum_256(a, s_red);
  b = block_sum_256(b, s_red);
  c = block_sum_256(c, s_red);
  if (threadIdx.x == 0) {
    out[i * 3 + 0] = a / (float)S;
    out[i * 3 + 1] = b / (float)S;
    out[i * 3 + 2] = c / (float)So;
  }
}

// out[j, c] += scale * sum_m G[m, j] * A[m, c]  (G NULL: ones, NV = 1): the thin weight-gradient products of the
// second-order backward (output layer 512 -> 1, bias sums).  CTA = slab of rows, thread = column(s), one atomicAdd per
// (CTA, column).
template <int NV>
__global__ void __launch_bounds__(256)
igr_colsums_kernel(const float* __restrict__ A, int64_t lda, int64_t M, int C, const float* __restrict__ G, int64_t ldg,
                   float scale, float* __restrict__ out, int64_t ldo, int rows_per_cta) {
  const int64_t m0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t m1 = min(M, m0 + rows_per_cta);
  for (int c = threadIdx.x; c < C; c += 256) {
    float acc[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] = 0.f;
    for (int64_t m = m0; m < m1; ++m) {
      const float x = __ldg(A + m * lda + c);
#pragma unroll
      for (int j = 0; j < NV; ++j) acc[j] = fmaf(G ? __ldg(G + m * ldg + j) : 1.f, x, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) atomicAdd(out + (size_t)j * ldo + c, acc[j] * scale);
  }
}

// delta_{L-1}[m, c] = ZE[m, c] + S[m, c] * f_bar[m] * w[c]: the output layer's data gradient (an outer product) with
// the extra pre-activation gradient of the reverse sweep's adjoint injected.  ZE or f_bar may be NULL.
__global__ void __launch_bounds__(256)
igr_seed_delta_kernel(const float* __restrict__ ZE, int64_t ldz, const float* __restrict__ S, int64_t lds,
                      const float* __restrict__ fbar, const float* __restrict__ w, int64_t M, int C,
                      float* __restrict__ out, int64_t ldo) {
  const int C4 = C >> 2;
  const int64_t total = M * C4;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = e / C4;
    const int c = (int)(e - m * C4) * 4;
    float4 r = ZE ? __ldg(reinterpret_cast<const float4*>(ZE + m * ldz + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (fbar) {
      const float fb = __ldg(fbar + m);
      const float4 s = __ldg(reinterpret_cast<const float4*>(S + m * lds + c));
      const float4 v = __ldg(reinterpret_cast<const float4*>(w + c));
      r.x = fmaf(s.x * fb, v.x, r.x); r.y = fmaf(s.y * fb, v.y, r.y);
      r.z = fmaf(s.z * fb, v.z, r.z); r.w = fmaf(s.w * fb, v.w, r.w);
    }
    *reinterpret_cast<float4*>(out + m * ldo + c) = r;
  }
}

// d latent[i, c] += scale * sum over the S rows of instance i of DX[r, c] (c < E);  d pts[r, :] (+)= scale * DX[r, E..E+1]
// (add_latent's backward: the latent code is repeated over an instance's points).  One CTA per instance.
__global__ void __launch_bounds__(256)
igr_latent_grad_kernel(const float* __restrict__ DX, int64_t lddx, int S, int E, float scale, float* __restrict__ dlat,
                       float* __restrict__ dpts, int accumulate_pts) {
  const int64_t i = blockIdx.x;
  for (int c = threadIdx.x; c < E + 2; c += 256) {
    if (c < E) {
      float acc = 0.f;
      for (int p = 0; p < S; ++p) acc += __ldg(DX + (i * S + p) * lddx + c);
      dlat[i * E + c] += acc * scale;
    } else if (dpts) {
      for (int p = 0; p < S; ++p) {
        const int64_t r = i * S + p;
        const float v = __ldg(DX + r * lddx + c) * scale;
        dpts[r * 2 + (c - E)] = accumulate_pts ? dpts[r * 2 + (c - E)] + v : v;
      }
    }
  }
}

// backward of igr_loss_terms_kernel: upstream d terms (I, 3) -> f_bar (I*S), g_bar on-surface (I*S, 2), g_bar
// off-surface (I*So, 2).  torch semantics: abs' = sign (0 at 0), ||.||_2' = v / ||v|| (0 at 0), min over
// {|g - n|, |g + n|} routes to the first on a tie (torch.min(dim) index).
__global__ void __launch_bounds__(256)
igr_loss_terms_bwd_kernel(const float* __restrict__ f_on, const float* __restrict__ g_on, const float* __restrict__ nrm,
                          const float* __restrict__ g_off, int S, int So, const float* __restrict__ dterms,
                          float* __restrict__ fbar, float* __restrict__ gbar_on, float* __restrict__ gbar_off) {
  const int64_t i = blockIdx.x;
  const float d0 = __ldg(dterms + i * 3 + 0) / (float)S, d1 = __ldg(dterms + i * 3 + 1) / (float)S;
  const float d2 = __ldg(dterms + i * 3 + 2) / (float)So;
  for (int p = threadIdx.x; p < S; p += 256) {
    const int64_t r = i * S + p;
    const float f = __ldg(f_on + r);
    fbar[r] = f > 0.f ? d0 : (f < 0.f ? -d0 : 0.f);
    const float gx = __ldg(g_on + 2 * r), gy = __ldg(g_on + 2 * r + 1);
    const float nx = __ldg(nrm + 2 * r), ny = __ldg(nrm + 2 * r + 1);
    const float sub = sqrtf((gx - nx) * (gx - nx) + (gy - ny) * (gy - ny));
    const float add = sqrtf((gx + nx) * (gx + nx) + (gy + ny) * (gy + ny));
    float ox = 0.f, oy = 0.f;
    if (sub <= add) {
      if (sub > 0.f) { ox = d1 * (gx - nx) / sub; oy = d1 * (gy - ny) / sub; }
    } else {
      if (add > 0.f) { ox = d1 * (gx + nx) / add; oy = d1 * (gy + ny) / add; }
    }
    gbar_on[2 * r] = ox;
    gbar_on[2 * r + 1] = oy;
  }
  for (int p = threadIdx.x; p < So; p += 256) {
    const int64_t r = i * So + p;
    const float gx = __ldg(g_off + 2 * r), gy = __ldg(g_off + 2 * r + 1);
    const float nn = sqrtf(gx * gx + gy * gy);
    const float k = nn > 0.f ? d2 * 2.f * (nn - 1.f) / nn : 0.f;
    gbar_off[2 * r] = k * gx;
    gbar_off[2 * r + 1] = k * gy;
  }
}

}  // namespace

extern "C" int p2c_igr_add_latent(const float* latent, const float* pts, int64_t R, int S, int E, float* X0, int64_t ldx,
                                  float* P, int64_t ldp, int colp, float pscale, void* stream) {
  if (!latent || !pts || !X0 || R <= 0 || S <= 0 || E <= 0 || ldx < E + 2 || (R % S) != 0) return P2C_EINVAL;
  if (P && ldp < colp + E + 2) return P2C_EINVAL;
  const int64_t total = R * ldx;
  const int blocks = (int)min((int64_t)148 * 16, (total + 255) / 256);
  igr_add_latent_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(latent, pts, R, S, E, X0, ldx, P, ldp, colp, pscale);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_igr_scale_cols(const float* A, int64_t lda, const float* w, int64_t M, int C, float* out,
                                  int64_t ldo, void* stream) {
  if (!A || !w || !out || M <= 0 || C <= 0 || lda < C || ldo < C) return P2C_EINVAL;
  if ((C % 4) || (lda % 4) || (ldo % 4) ||
      ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out)) & 15))
    return P2C_EALIGN;
  const int64_t total = M * (C / 4);
  const int blocks = (int)min((int64_t)148 * 16, (total + 255) / 256);
  igr_scale_cols_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(A, lda, w, M, C, out, ldo);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_igr_rowdots(const float* A, int64_t lda, int64_t M, int C, const float* V, int64_t v_stride_j,
                               int64_t v_stride_c, int NV, const float* bias, float scale, float* out, int64_t ldo,
                               int accumulate, void* stream) {
  if (!A || !V || !out || M <= 0 || C <= 0 || lda < C || ldo < NV) return P2C_EINVAL;
  if (NV != 1 && NV != 2) return P2C_EUNSUPPORTED;
  if ((C % 4) || (lda % 4) || (reinterpret_cast<uintptr_t>(A) & 15)) return P2C_EALIGN;
  const size_t smem = (size_t)NV * C * sizeof(float);
  if (smem > 48 * 1024) return P2C_EUNSUPPORTED;
  const int blocks = (int)min((int64_t)148 * 8, (M + 7) / 8);
  if (NV == 1)
    igr_rowdots_kernel<1><<<blocks, 256, smem, (cudaStream_t)stream>>>(A, lda, M, C, V, v_stride_j, v_stride_c, bias, scale, out,
                                                                      ldo, accumulate);
  else
    igr_rowdots_kernel<2><<<blocks, 256, smem, (cudaStream_t)stream>>>(A, lda, M, C, V, v_stride_j, v_stride_c, bias, scale, out,
                                                                      ldo, accumulate);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_igr_loss_terms(const float* f_on, const float* g_on, const float* normals, const float* g_off,
                                  int instances, int S, int S_off, float* out, void* stream) {
  if (!f_on || !g_on || !normals || !g_off || !out || instances <= 0 || S <= 0 || S_off <= 0) return P2C_EINVAL;
  igr_loss_terms_kernel<<<instances, 256, 0, (cudaStream_t)stream>>>(f_on, g_on, normals, g_off, S, S_off, out);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_igr_colsums(const float* A, int64_t lda, int64_t M, int C, const float* G, int64_t ldg, int NV,
                               float scale, float* out, int64_t ldo, void* stream) {
  if (!A || !out || M <= 0 || C <= 0 || lda < C || ldo < C) return P2C_EINVAL;
  if ((NV != 1 && NV != 2) || (!G && NV != 1) || (G && ldg < NV)) return P2C_EINVAL;
  int64_t ctas = min((int64_t)148 * 8, (M + 63) / 64);
  const int rows = (int)((M + ctas - 1) / ctas);
  ctas = (M + rows - 1) / rows;
  if (NV == 1) igr_colsums_kernel<1><<<(int)ctas, 256, 0, (cudaStream_t)stream>>>(A, lda, M, C, G, ldg, scale, out, ldo, rows);
  else igr_colsums_kernel<2><<<(int)ctas, 256, 0, (cudaStream_t)stream>>>(A, lda, M, C, G, ldg, scale, out, ldo, rows);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_igr_seed_delta(const float* ZE, int64_t ldz, const float* S, int64_t lds, const float* f_bar,
                                  const float* w, int64_t M, int C, float* out, int64_t ldo, void* stream) {
  if (!out || M <= 0 || C <= 0 || ldo < C || (!ZE && !f_bar)) return P2C_EINVAL;
  if ((ZE && ldz < C) || (f_bar && (!S || !w || lds < C))) return P2C_EINVAL;
  if ((C % 4) || (ldo % 4) || (ZE && (ldz % 4)) || (f_bar && (lds % 4)) ||
      ((reinterpret_cast<uintptr_t>(ZE) | reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(w) |
        reinterpret_cast<uintptr_t>(out)) & 15))
    return P2C_EALIGN;
  const int64_t total = M * (C / 4);
  const int blocks = (int)min((int64_t)148 * 16, (total + 255) / 256);
  igr_seed_delta_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ZE, ldz, S, lds, f_bar, w, M, C, out, ldo);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_igr_latent_grad(const float* DX, int64_t lddx, int instances, int S, int E, float scale, float* dlatent,
                                   float* dpts, int accumulate_pts, void* stream) {
  if (!DX || !dlatent || instances <= 0 || S <= 0 || E <= 0 || lddx < E + 2) return P2C_EINVAL;
  igr_latent_grad_kernel<<<instances, 256, 0, (cudaStream_t)stream>>>(DX, lddx, S, E, scale, dlatent, dpts, accumulate_pts);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_igr_loss_terms_bwd(const float* f_on, const float* g_on, const float* normals, const float* g_off,
                                      int instances, int S, int S_off, const float* dterms, float* f_bar, float* g_bar_on,
                                      float* g_bar_off, void* stream) {
  if (!f_on || !g_on || !normals || !g_off || !dterms || !f_bar || !g_bar_on || !g_bar_off || instances <= 0 || S <= 0 ||
      S_off <= 0)
    return P2C_EINVAL;
  igr_loss_terms_bwd_kernel<<<instances, 256, 0, (cudaStream_t)stream>>>(f_on, g_on, normals, g_off, S, S_off, dterms, f_bar,
                                                                        g_bar_on, g_bar_off);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
