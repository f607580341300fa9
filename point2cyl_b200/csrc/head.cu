// Output heads of the backbone (replaces dropout + fc2 at models/pointnet_extrusion.py:60-65):
//   Y[m, j] = sum_k relu(H[m,k]*scale[k]+shift[k]) * mask[b, k, n] * W[j, k] + bias[j],   m = b*N + n
// H is fc1's raw output (rows, C); scale/shift fold bn1; mask is the (B, C, N) channel-first tensor
// F.dropout(ones) returns - the reference's own layout, so no transpose copy is needed: with one thread per
// point, mask[b, k, n..n+31] is a coalesced 128-byte read for every k.  The few output channels (3 + 2K <= 36)
// make this an HBM-bound SIMT kernel (reads H and the mask once: 8*C bytes per point), not a tensor-core one.
// Each thread streams its own H row 64 bytes (two full sectors) at a time, software-pipelined one step ahead;
// W^T lives in shared memory and is read as broadcast float4; the Y tile is staged through shared memory so
// the (rows x Nout) block leaves as one contiguous run.
#include "bn_fold.cuh"

namespace {

constexpr int HEAD_ROWS = 256;   // threads = points per CTA
constexpr int KC = 16;           // channels per software-pipeline step (64 B = two full sectors of a row)

template <int NP>   // NP = Nout padded to a multiple of 4
__global__ void __launch_bounds__(HEAD_ROWS)
head_kernel(const float* __restrict__ H, int64_t ldh, const float* __restrict__ scale, const float* __restrict__ shift,
            const float* __restrict__ mask_cf, const int64_t* __restrict__ seed, const float* __restrict__ W,
            const float* __restrict__ bias, float* __restrict__ Y, int64_t ldy, int64_t M, int N, int C, int Nout,
            const BnFoldDev bn) {
  extern __shared__ __align__(16) float sm[];
  float* s_wt = sm;                           // [C][NP]   (W transposed, zero padded)
  float* s_sc = s_wt + C * NP;                // [C]
  float* s_sh = s_sc + C;                     // [C]
  float* s_y = s_sh + C;                      // [HEAD_ROWS][Nout] output staging
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * HEAD_ROWS;

  for (int e = tid; e < C * NP; e += HEAD_ROWS) {
    const int k = e / NP, j = e - k * NP;
    s_wt[e] = j < Nout ? __ldg(W + (size_t)j * C + k) : 0.f;
  }
  for (int k = tid; k < C; k += HEAD_ROWS) {
    if (bn.active) {                          // pending bn1: folded here, CTA 0 publishes it
      p2c_bn_fold_channel(bn, k, blockIdx.x == 0, s_sc[k], s_sh[k]);
    } else {
      s_sc[k] = scale ? __ldg(scale + k) : 1.f;
      s_sh[k] = shift ? __ldg(shift + k) : 0.f;
    }
  }
  __syncthreads();
  const bool affine = scale != nullptr || bn.active;

  const int64_t m = m0 + tid;
  const bool ok = m < M;
  const int64_t mm = ok ? m : M - 1;          // clamp: out-of-range threads compute a duplicate row, never store
  const int64_t b = mm / N;
  const int n = (int)(mm - b * N);
  const float* mk = mask_cf ? mask_cf + (size_t)b * C * N + n : nullptr;
  const float* hrow = H + (size_t)mm * ldh;
  float acc[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) acc[j] = 0.f;

  // one thread per point: its H row arrives as full sectors (16 channels = 64 B per step), the mask column
  // mask[b, k, n] is a coalesced 128-byte warp read for every k; loads of step c+1 are issued before the
  // FMAs of step c
  float4 hq[KC / 4];
  float mq[KC];
  P2CPhilox4 bits;                            // in-kernel dropout: 128 keep-bits of this point per Philox call
  auto load_step = [&](int k0) {
#pragma unroll
    for (int v = 0; v < KC / 4; ++v) hq[v] = __ldg(reinterpret_cast<const float4*>(hrow + k0) + v);
    if (seed && !mk) {
      if ((k0 & 127) == 0) bits = p2c_dropout_bits(seed, mm, k0 >> 7);
#pragma unroll
      for (int i = 0; i < KC; ++i) mq[i] = p2c_dropout_scale(bits, k0 + i);
    } else {
#pragma unroll
      for (int i = 0; i < KC; ++i) mq[i] = mk ? __ldg(mk + (size_t)(k0 + i) * N) : 1.f;
    }
  };
  load_step(0);
  for (int k0 = 0; k0 < C; k0 += KC) {
    float x[KC];
#pragma unroll
    for (int v = 0; v < KC / 4; ++v) { x[v * 4] = hq[v].x; x[v * 4 + 1] = hq[v].y; x[v * 4 + 2] = hq[v].z; x[v * 4 + 3] = hq[v].w; }
    float mcur[KC];
#pragma unroll
    for (int i = 0; i < KC; ++i) mcur[i] = mq[i];
    if (k0 + KC < C) load_step(k0 + KC);
#pragma unroll
    for (int i = 0; i < KC; ++i) {
      float xi = x[i];
      if (affine) xi = fmaxf(fmaf(xi, s_sc[k0 + i], s_sh[k0 + i]), 0.f);
      xi *= mcur[i];
      const float* wr = s_wt + (k0 + i) * NP;
#pragma unroll
      for (int j = 0; j < NP; j += 4) {
        const float4 w = *reinterpret_cast<const float4*>(wr + j);
        acc[j] = fmaf(xi, w.x, acc[j]);
        acc[j + 1] = fmaf(xi, w.y, acc[j + 1]);
        acc[j + 2] = fmaf(xi, w.z, acc[j + 2]);
        acc[j + 3] = fmaf(xi, w.w, acc[j + 3]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NP; ++j)
    if (j < Nout) s_y[tid * Nout + j] = acc[j] + (bias ? __ldg(bias + j) : 0.f);
  __syncthreads();
  const int rows = (int)min((int64_t)HEAD_ROWS, M - m0);
  if (ldy == Nout) {                            // contiguous block of rows*Nout floats
    float* yo = Y + (size_t)m0 * Nout;
    for (int e = tid; e < rows * Nout; e += HEAD_ROWS) yo[e] = s_y[e];
  } else {
    for (int e = tid; e < rows * Nout; e += HEAD_ROWS) {
      const int r = e / Nout, j = e - r * Nout;
      Y[(size_t)(m0 + r) * ldy + j] = s_y[e];
    }
  }
}

template <int NP>
int launch_head(const float* H, int64_t ldh, const float* scale, const float* shift, const float* mask_cf,
                const int64_t* seed, const float* W, const float* bias, float* Y, int64_t ldy, int64_t M, int N, int C,
                int Nout, const BnFoldDev& bn, cudaStream_t st) {
  const size_t smem = ((size_t)C * NP + 2 * C + (size_t)HEAD_ROWS * Nout) * sizeof(float);
  auto k = head_kernel<NP>;
  if (smem > 48 * 1024) P2C_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<p2c_ceil_div(M, HEAD_ROWS), HEAD_ROWS, smem, st>>>(H, ldh, scale, shift, mask_cf, seed, W, bias, Y, ldy, M, N, C, Nout, bn);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

}  // namespace

// linear_tc.cu: the tcgen05 layer kernel; drop_seed makes its operand transform draw the dropout mask
int p2c_linear_tc(const float* X, int64_t ldx, const float* W, const float* bias, const float* in_scale,
                  const float* in_shift, const float* in_mask, int64_t ldmask, float* Y, int64_t ldy, int M, int N,
                  int K, double* stats, int pool_group, float* Ymax, float* Ymin, int precision,
                  const p2c_bn_fold* in_bn, cudaStream_t st, const int64_t* drop_seed, const P2cXyzFirst* xyz_first);

extern "C" int p2c_head_masked(const float* H, int64_t ldh, const float* scale, const float* shift,
                               const float* mask_cf, const int64_t* dropout_seed, const float* W, const float* bias,
                               float* Y, int64_t ldy, int B, int N, int C, int Nout, const p2c_bn_fold* bn,
                               int precision, void* stream) {
  if (!H || !W || !Y || B <= 0 || N <= 0 || C <= 0 || Nout <= 0 || ldh < C || ldy < Nout) return P2C_EINVAL;
  if ((scale == nullptr) != (shift == nullptr)) return P2C_EINVAL;
  if (int e = p2c_bn_fold_check(bn, C)) return e;
  if (precision == P2C_PREC_3XTF32 && !mask_cf && (int64_t)B * N < (1ll << 31) && (scale || bn)) {
    // The heads are one more per-point layer: relu(bn1(fc1)) * dropout -> (3 + 2K) channels.  On the tcgen05 layer
    // kernel (weights resident in tensor memory, the mask drawn in the operand transform from the same Philox counter
    // this file and p2c_head_bwd use) the launch is bound by reading H once; the SIMT kernel below - one thread per
    // point, 2560 FMAs and 640 shared-memory loads each - ran at a fifth of that.
    const int rc = p2c_linear_tc(H, ldh, W, bias, scale, shift, nullptr, 0, Y, ldy, B * N, Nout, C, nullptr, 0, nullptr,
                                 nullptr, precision, bn, (cudaStream_t)stream, dropout_seed, nullptr);
    if (rc != P2C_EUNSUPPORTED) return rc;
  }
  const BnFoldDev bnd = p2c_bn_fold_dev(bn);
  if (C % 16 != 0 || (ldh % 4) != 0 || (reinterpret_cast<uintptr_t>(H) & 15) != 0) return P2C_EALIGN;
  if (C > 256 || Nout > 36) return P2C_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t M = (int64_t)B * N;
#define P2C_HEAD(NPV) return launch_head<NPV>(H, ldh, scale, shift, mask_cf, dropout_seed, W, bias, Y, ldy, M, N, C, Nout, bnd, st)
  if (Nout <= 4) P2C_HEAD(4);
  if (Nout <= 8) P2C_HEAD(8);
  if (Nout <= 12) P2C_HEAD(12);
  if (Nout <= 20) P2C_HEAD(20);
  if (Nout <= 28) P2C_HEAD(28);
  P2C_HEAD(36);
#undef P2C_HEAD
}
