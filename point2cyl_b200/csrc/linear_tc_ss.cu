// tcgen05 MLP layer for LARGE K (K > 192: sa3, fp3, fp2 of the backbone), where the weight matrix no longer
// fits tensor memory.  Same transposed formulation and epilogue as linear_tc.cu (D[n, m] = sum_k W[n,k] f(X)[m,k],
// one output channel per TMEM lane), but both UMMA operands come from shared memory:
//   A = W_hi / W_lo k-blocks [128 channels x 32 floats], TMA-loaded from a pre-split copy of the weights
//       (p2c_split_tf32: hi = low 13 mantissa bits cleared, lo = w - hi, rows padded to 16 bytes for TMA);
//   B = transformed activation k-blocks (hi | lo), as in linear_tc.cu.
// These layers have few rows (B*128 or B*512), so the work list is (row tile, channel tile) pairs handed out
// round-robin to a persistent grid; consecutive CTAs share a row tile (L2 reuse of X).
#include <cstdlib>

#include "bn_fold.cuh"
#include "tc_common.cuh"

using namespace p2c_tc;

namespace {

constexpr int SS_THREADS = 512;   // 16 warps: producer, MMA, alloc, idle, 4 epilogue, 4 transform, 4 epilogue (second group)
constexpr int SS_MAX_RAW = 4, SS_MAX_XT = 3;

struct SsArgs {
  const float* bias;
  const float* in_scale; const float* in_shift;
  float* Y; int64_t ldy;
  int M, N, K, KB;
  double* stats;
  int pool_group;
  float* Ymax; float* Ymin;
  int raw_stages, xt_stages;
  int m_tiles, n_tiles;
  int y_tma;
  BnFoldDev bn;     // pending BatchNorm of X, folded in the prologue (bn.active) instead of in_scale / in_shift
  // activation / multiplier epilogues of the implicit-network layers (template parameter EPI, p2c_linear_act):
  //   EPI 1: Y = softplus_beta(acc + bias) * oscale, S = sigmoid(beta (acc + bias));  EPI 2: Y = (acc + bias) * Mul * oscale
  float beta, oscale;
  int s_tma;                     // the second output S leaves through TMA stores too
  const float* mul; int64_t ldmul;
  float* S; int64_t lds;
  // second-order backward of the implicit network (p2c_linear_act_bwd), no bias:
  //   EPI 3: Y = acc * Mul * oscale, S = beta * acc * mul2 * (1 - Mul)   (adjoint of the reverse sweep: Mul = softplus',
  //          mul2 = a_i, so S is the extra pre-activation gradient a_bar * q * softplus'')
  //   EPI 4: Y = acc * Mul * oscale + mul2                              (data gradient with that extra term injected)
  const float* mul2; int64_t ldmul2;
  int raw_hi;                    // RAW tile = hi operand, XT ring holds lo only (no operand transform needed)
  const float* bias_rows; int bias_group;   // EPI 0: bias[(row / bias_group), n] (N floats per group of rows) in place of bias[n]
  int dbg_mode;                  // tools only (env P2C_SS_DBG): 1 skip the epilogue body, 2 skip the correction read,
                                 // 4 skip the correction MMAs, 8 transform copies hi only (timing experiments)
};

// ---- bf16 mode (P2C_PREC_BF16): one kind::f16 MMA pass on bf16 operands instead of the three tf32 passes ----
// The fp32 activations still arrive as [128 rows x 32 floats] TMA boxes; the transform warps apply the BatchNorm+ReLU
// fold and write a bf16 tile [128 rows x 32 bf16 = 64 B] in the K-major SWIZZLE_64B layout (16-byte chunk index XOR
// (row >> 1) & 3); the weights come from a bf16 copy (p2c_cast_bf16) through a bf16 TMA map with the same swizzle.
// Two MMAs (K = 16) per k-block, fp32 accumulate; epilogue unchanged.  Not fp32-faithful (8-bit mantissa): it exists
// for BASELINE.json's bf16 MLP-stack configuration, not for the parity path.
__device__ __forceinline__ uint64_t make_kmajor_sw64_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;         // 8 rows x 64 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                  // SWIZZLE_64B
  return d;
}
constexpr uint32_t TC_IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BM >> 3) << 17) |
                                   ((uint32_t)(TC_BN >> 4) << 24);
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
constexpr int W_BF16_BYTES = TC_BN * TC_BK * 2;   // 8 KB

__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct SsSmem {
  uint32_t raw_off, xt_off, w_off, ystage_off, scale_off, shift_off, bar_off, total;
};
// raw_hi: the activation needs no operand transform (no BatchNorm fold), so the RAW fp32 tile itself is the tf32 "hi"
// operand (the tensor core ignores the low 13 mantissa bits) and the XT ring holds only the lo tiles: 64 KB instead
// of 80 KB per pipeline stage, i.e. three stages in flight instead of two for K = 512.
__host__ __device__ inline SsSmem ss_smem_layout(int KB, int raw_stages, int xt_stages, int y_stage, int raw_hi = 0) {
  SsSmem L;
  uint32_t o = 0;
  L.raw_off = o;    o += (uint32_t)raw_stages * RAW_BYTES;
  L.xt_off = o;     o += (uint32_t)xt_stages * (raw_hi ? 1u : 2u) * RAW_BYTES;   // [stage][hi|lo] or [stage][lo]
  L.w_off = o;      o += (uint32_t)xt_stages * 2u * RAW_BYTES;   // [stage][hi|lo]
  L.ystage_off = o; o += y_stage ? 4u * 2u * 4096u : 0u;        // one 4 KB staging tile per epilogue warp
  L.scale_off = o;  o += raw_hi ? 0u : (uint32_t)KB * TC_BK * 4u;
  L.shift_off = o;  o += raw_hi ? 0u : (uint32_t)KB * TC_BK * 4u;
  L.bar_off = o;    o += 512u;
  L.total = o;
  return L;
}

template <bool BF16, int EPI>
__global__ void __launch_bounds__(SS_THREADS, 1)
linear_tc_ss_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWhi,
                    const __grid_constant__ CUtensorMap tmWlo, const __grid_constant__ CUtensorMap tmY,
                    const __grid_constant__ CUtensorMap tmS, const SsArgs a) {
#ifdef P2C_SS_DEBUG      // timing toggles (tools/igr_exp.sh): build with NVCC_FLAGS=-DP2C_SS_DEBUG; compiled out otherwise
  const int dbg_mode = a.dbg_mode;
#else
  constexpr int dbg_mode = 0;
#endif
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const SsSmem L = ss_smem_layout(a.KB, a.raw_stages, a.xt_stages, a.y_tma, a.raw_hi);
  const bool raw_hi = !BF16 && a.raw_hi;
  const uint32_t xt_stage_bytes = raw_hi ? RAW_BYTES : 2u * RAW_BYTES;
  uint8_t* raw_sm = smem + L.raw_off;
  uint8_t* xt_sm = smem + L.xt_off;
  uint8_t* w_sm = smem + L.w_off;
  uint8_t* ystage = smem + L.ystage_off;
  float* s_scale = reinterpret_cast<float*>(smem + L.scale_off);
  float* s_shift = reinterpret_cast<float*>(smem + L.shift_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* raw_full = bars;                          // [SS_MAX_RAW]
  uint64_t* raw_empty = raw_full + SS_MAX_RAW;
  uint64_t* xt_full = raw_empty + SS_MAX_RAW;         // [SS_MAX_XT]
  uint64_t* xt_empty = xt_full + SS_MAX_XT;
  uint64_t* w_full = xt_empty + SS_MAX_XT;            // [SS_MAX_XT]
  uint64_t* w_empty = w_full + SS_MAX_XT;
  uint64_t* acc_full = w_empty + SS_MAX_XT;           // [2]
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int KB = a.KB, KPAD = a.KB * TC_BK;
  const int RS = a.raw_stages, XS = a.xt_stages;
  const int total_tiles = a.m_tiles * a.n_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmWhi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmWlo)) : "memory");
    // raw_hi: a RAW stage is released by the MMAs that read it (one tcgen05.commit), else by the 4 transform warps
    for (int s = 0; s < RS; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], raw_hi ? 1 : 4); }
    for (int s = 0; s < XS; ++s) {
      mbar_init(&xt_full[s], 4); mbar_init(&xt_empty[s], 1);
      mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1);
    }
    // EPI != 0: both epilogue groups drain every tile (two 32-row chunks each), so 8 warps release an accumulator
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], EPI != 0 ? 8 : 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);      // 2 x (main accumulator | correction accumulator), 128 columns each
  p2c_grid_dep_wait();          // (common.cuh) above: CTA-local setup; below: data of predecessor kernels
  p2c_grid_dep_launch();
  for (int k = tid; !raw_hi && k < KPAD; k += SS_THREADS) {
    float sc = 0.f, sh = 0.f;
    if (k < a.K) {
      if (a.bn.active) p2c_bn_fold_channel(a.bn, k, blockIdx.x == 0, sc, sh);
      else if (a.in_scale) { sc = __ldg(a.in_scale + k); sh = __ldg(a.in_shift + k); }
    }
    s_scale[k] = sc;
    s_shift[k] = sh;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto tile_mt = [&](int t) { return ((int)blockIdx.x + t * (int)gridDim.x) / a.n_tiles; };
  auto tile_nt = [&](int t) { return ((int)blockIdx.x + t * (int)gridDim.x) % a.n_tiles; };

  if (warp == 0) {
    // ===== TMA producer: X k-block -> RAW ring, W_hi / W_lo k-blocks -> W ring (warp-uniform loop, elected issuer) =====
    {
      int s = 0; uint32_t ph = 0;
      int ws = 0; uint32_t wph = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int m0 = tile_mt(t) * TC_BM, n0 = tile_nt(t) * TC_BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&w_empty[ws], wph ^ 1);
          if (elect_one_sync()) {
            if (BF16) {
              mbar_arrive_expect_tx(&w_full[ws], W_BF16_BYTES);
              tma_load_2d(w_sm + (size_t)ws * 2 * RAW_BYTES, &tmWhi, &w_full[ws], kb * TC_BK, n0);
            } else {
              mbar_arrive_expect_tx(&w_full[ws], 2 * RAW_BYTES);
              tma_load_2d(w_sm + (size_t)ws * 2 * RAW_BYTES, &tmWhi, &w_full[ws], kb * TC_BK, n0);
              tma_load_2d(w_sm + (size_t)ws * 2 * RAW_BYTES + RAW_BYTES, &tmWlo, &w_full[ws], kb * TC_BK, n0);
            }
          }
          __syncwarp();
          if (++ws == XS) { ws = 0; wph ^= 1; }
          mbar_wait(&raw_empty[s], ph ^ 1);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&raw_full[s], RAW_BYTES);
            tma_load_2d(raw_sm + (size_t)s * RAW_BYTES, &tmX, &raw_full[s], kb * TC_BK, m0);
          }
          __syncwarp();
          if (++s == RS) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-uniform loop, one elected lane issues (see tc_common.cuh: elect_one_sync) =====
    {
      int xs = 0; uint32_t xph = 0;
      int rs = 0;                                      // RAW stage of this k-block (raw_hi mode: the hi operand)
      const bool corr = !(dbg_mode & 4);
      for (int t = 0; t < my_tiles; ++t) {
        const int ab = t & 1;
        const uint32_t accph = (uint32_t)(t >> 1) & 1u;
        mbar_wait(&acc_empty[ab], accph ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)ab * TC_BM;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&w_full[xs], xph);
          mbar_wait(&xt_full[xs], xph);
          tc_fence_after();
          const uint32_t xt0 = smem_u32(xt_sm + (size_t)xs * xt_stage_bytes);
          const uint32_t x_hi = raw_hi ? smem_u32(raw_sm + (size_t)rs * RAW_BYTES) : xt0;
          const uint32_t x_lo = raw_hi ? xt0 : xt0 + RAW_BYTES;
          const uint32_t w_hi = smem_u32(w_sm + (size_t)xs * 2 * RAW_BYTES), w_lo = w_hi + RAW_BYTES;
          if (elect_one_sync()) {
            if (BF16) {
              const uint64_t a0 = make_kmajor_sw64_desc(w_hi), b0 = make_kmajor_sw64_desc(x_hi);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)           // 16 bf16 = 32 bytes per k-step
                umma_bf16_ss(d, a0 + (uint64_t)(ks * 2), b0 + (uint64_t)(ks * 2), TC_IDESC_BF16, (kb | ks) != 0);
            } else {
              // The tensor core's fp32 accumulate TRUNCATES (measured: a relative shrink of ~2.3e-8 per accumulate
              // step, tests/tools/precision_probe.py).  The two correction products are 2^-11 of the main one, so
              // they go to their own accumulator (columns 256..511) where that truncation is negligible, and the main
              // accumulator sees K/8 accumulate steps instead of 3K/8; the epilogue adds the two in round-to-nearest
              // fp32.  (Four main products back to back, then the eight corrections: the tensor pipe chains
              // accumulations into one accumulator better than it alternates between two.)
              // The four 8-column k-steps of a tile differ in the descriptor's start-address field only (+32 B = +2).
              const uint64_t ahi = make_kmajor_sw128_desc(w_hi), alo = make_kmajor_sw128_desc(w_lo);
              const uint64_t bhi = make_kmajor_sw128_desc(x_hi), blo = make_kmajor_sw128_desc(x_lo);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_tf32_ss(d, ahi + (uint64_t)(ks * 2), bhi + (uint64_t)(ks * 2), TC_IDESC, (kb | ks) != 0);
              if (corr) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  umma_tf32_ss(d + 256u, alo + (uint64_t)(ks * 2), bhi + (uint64_t)(ks * 2), TC_IDESC, (kb | ks) != 0);
                  umma_tf32_ss(d + 256u, ahi + (uint64_t)(ks * 2), blo + (uint64_t)(ks * 2), TC_IDESC, 1u);
                }
              }
            }
            umma_commit(&xt_empty[xs]);
            umma_commit(&w_empty[xs]);
            if (raw_hi) umma_commit(&raw_empty[rs]);     // the RAW tile was an MMA operand: free it when they retire
            if (kb == KB - 1) umma_commit(&acc_full[ab]);
          }
          __syncwarp();
          if (++xs == XS) { xs = 0; xph ^= 1; }
          if (++rs == RS) rs = 0;
        }
      }
    }
  } else if (warp >= 8 && warp < 12) {
    // ===== operand transform (see linear_tc.cu: thread = (16-byte chunk cj, row mod 8, half), its eight rows 1024 B
    // apart, so the swizzled position is one per-thread constant and every access is [base + immediate]) =====
    const int tt = tid - 256;
    const int cj = tt & 7, r7 = (tt >> 3) & 7, hf = tt >> 6;
    const uint32_t toff = (uint32_t)(hf * 8192 + r7 * 128 + ((cj ^ r7) << 4));
    // bf16 tile: rows of 64 bytes, SWIZZLE_64B (16-byte chunk cj >> 1 XOR (row >> 1) & 3), rows 8 apart = 512 B
    const uint32_t toff16 = (uint32_t)((hf * 64 + r7) * 64 + ((((unsigned)cj >> 1) ^ (((unsigned)r7 >> 1) & 3u)) << 4) + ((cj & 1) << 3));
    const bool has_affine = !raw_hi && (a.in_scale != nullptr || a.bn.active);
    int s = 0; uint32_t ph = 0;
    int xs = 0; uint32_t xph = 0;
    for (int t = 0; t < my_tiles; ++t) {
      for (int kb = 0; kb < KB; ++kb) {
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_affine) {
          sc = *reinterpret_cast<const float4*>(s_scale + kb * TC_BK + cj * 4);
          sh = *reinterpret_cast<const float4*>(s_shift + kb * TC_BK + cj * 4);
        }
        mbar_wait(&raw_full[s], ph);
        const uint8_t* rawp = raw_sm + (size_t)s * RAW_BYTES + toff;
        float4 x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = *reinterpret_cast<const float4*>(rawp + i * 1024);
        __syncwarp();
        if (lane == 0 && !raw_hi) mbar_arrive(&raw_empty[s]);
        if (++s == RS) { s = 0; ph ^= 1; }
        if (has_affine) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 p0 = __ffma2_rn(make_float2(x[i].x, x[i].y), make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
            const float2 p1 = __ffma2_rn(make_float2(x[i].z, x[i].w), make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
            x[i] = make_float4(fmaxf(p0.x, 0.f), fmaxf(p0.y, 0.f), fmaxf(p1.x, 0.f), fmaxf(p1.y, 0.f));
          }
        }
        mbar_wait(&xt_empty[xs], xph ^ 1);
        uint8_t* hip = xt_sm + (size_t)xs * xt_stage_bytes;
        if (BF16) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<uint2*>(hip + toff16 + i * 512) = make_uint2(pack_bf16(x[i].x, x[i].y), pack_bf16(x[i].z, x[i].w));
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 h;
            h.x = __uint_as_float(__float_as_uint(x[i].x) & 0xffffe000u);
            h.y = __uint_as_float(__float_as_uint(x[i].y) & 0xffffe000u);
            h.z = __uint_as_float(__float_as_uint(x[i].z) & 0xffffe000u);
            h.w = __uint_as_float(__float_as_uint(x[i].w) & 0xffffe000u);
            const float2 l0 = __fadd2_rn(make_float2(x[i].x, x[i].y), make_float2(-h.x, -h.y));
            const float2 l1 = __fadd2_rn(make_float2(x[i].z, x[i].w), make_float2(-h.z, -h.w));
            const float4 l = make_float4(l0.x, l0.y, l1.x, l1.y);
            if (raw_hi) {                                // lo tile only: the RAW tile itself is the hi operand
              *reinterpret_cast<float4*>(hip + toff + i * 1024) = l;
            } else {
              *reinterpret_cast<float4*>(hip + toff + i * 1024) = h;
              *reinterpret_cast<float4*>(hip + RAW_BYTES + toff + i * 1024) = l;
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&xt_full[xs]);
        if (++xs == XS) { xs = 0; xph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread = output channel of the tile; two groups of four warps (4-7, 12-15) drain the two
    // accumulator buffers alternately, like linear_tc.cu =====
    const int grp = warp >= 12 ? 1 : 0;
    const int q = warp & 3;
    const int ch = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float NEG_INF = -__int_as_float(0x7f800000), POS_INF = __int_as_float(0x7f800000);
    const int G = a.pool_group;
    const int gshift = G ? 31 - __clz(G) : 0;
    float* ystg = reinterpret_cast<float*>(ystage + (size_t)(grp * 4 + q) * 4096);   // one 4 KB staging tile per warp
    const bool y_tma_all = a.Y != nullptr && a.y_tma;
    // EPI == 0: group g drains the tiles t = g (mod 2) whole (the BatchNorm / pool reductions run over a tile's rows).
    // EPI != 0 (element-wise epilogues): BOTH groups drain every tile, chunks {0,1} and {2,3} - an accumulator is held
    // for a quarter of a tile's epilogue instead of three quarters, which is what lets the MMAs of tile t+2 start on
    // time (measured: the multiplier epilogue's dependent global loads made the old scheme epilogue-bound).
    const int t_step = EPI != 0 ? 1 : 2;
    const int c_lo = EPI != 0 ? 2 * grp : 0, c_hi = EPI != 0 ? 2 * grp + 2 : 4;
    for (int t = (EPI != 0 ? 0 : grp); t < my_tiles; t += t_step) {
      const int ab = t & 1;
      const uint32_t accph = (uint32_t)(t >> 1) & 1u;
      const int m0 = tile_mt(t) * TC_BM, n0 = tile_nt(t) * TC_BN;
      const int n = n0 + ch;
      const bool n_ok = n < a.N;
      if (EPI >= 2 && n_ok) {
        // the multiplier rows this warp will read in its two chunks: pull them into L2 while the MMAs of the tile run
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int row = m0 + c_lo * 32 + j * 32 + lane;
          if (row < a.M) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.mul + (size_t)row * a.ldmul + n0 + q * 32));
            if (EPI >= 3) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.mul2 + (size_t)row * a.ldmul2 + n0 + q * 32));
          }
        }
      }
      // A [32 x 32] box that sticks out over channel N is NOT left to the TMA unit (measured: its stores clip at
      // 16-byte granularity, so with N % 4 != 0 up to three channels beyond N were written): this warp then stores its
      // rows itself - lanes = consecutive channels, so each store instruction is still one coalesced line.
      const bool y_tma = y_tma_all && (n0 + q * 32 + 32 <= a.N);
      const float bias = (a.bias && n_ok) ? __ldg(a.bias + n) : 0.f;
      float gmx = NEG_INF, gmn = POS_INF;
      float t1 = 0.f, t2 = 0.f;
      mbar_wait(&acc_full[ab], accph);
      tc_fence_after();
#pragma unroll 1                                       // rolled on purpose: instruction-cache footprint (see linear_tc.cu)
      for (int c = c_lo; c < c_hi; ++c) {
        uint32_t raw[32];
        tmem_ld32(tmem_base + (uint32_t)ab * TC_BM + (uint32_t)c * 32u + lane_addr, raw);
        if (!BF16 && !(dbg_mode & 2)) {              // + the correction accumulator (see the MMA issuer)
          uint32_t cor[32];
          tmem_ld32(tmem_base + 256u + (uint32_t)ab * TC_BM + (uint32_t)c * 32u + lane_addr, cor);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(__uint_as_float(raw[j]) + __uint_as_float(cor[j]));
        }
        tmem_wait_ld();
        if (c == c_hi - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[ab]);
        }
        const int mrow = m0 + c * 32;
        const int jmax = min(32, a.M - mrow);
        if (jmax <= 0 || (dbg_mode & 1)) continue;
        float* st = ystg;
        if (y_tma) {
          if (elect_one_sync()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous store has read it
          __syncwarp();
        }
        if (EPI != 0 && !y_tma) {
          // a box that is partial in the channel direction: guarded direct stores (see above)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (!(n_ok && j < jmax)) continue;
            const float z = __uint_as_float(raw[j]) + bias;
            const size_t row = (size_t)(mrow + j);
            if (EPI == 1) {
              const float bz = a.beta * z;
              float h = z, sg = 1.f;
              if (!(bz > 20.f)) {
                const float e = expf(bz);
                h = __fdiv_rn(log1pf(e), a.beta);
                sg = __fdiv_rn(e, 1.f + e);
              }
              a.Y[row * a.ldy + n] = h * a.oscale;
              if (a.s_tma) a.S[row * a.lds + n] = sg;
            } else if (EPI == 2) {
              a.Y[row * a.ldy + n] = z * __ldg(a.mul + row * a.ldmul + n) * a.oscale;
            } else {
              const float m = __ldg(a.mul + row * a.ldmul + n), v = __ldg(a.mul2 + row * a.ldmul2 + n);
              if (EPI == 3) {
                a.Y[row * a.ldy + n] = z * m * a.oscale;
                a.S[row * a.lds + n] = a.beta * z * v * (1.f - m);
              } else {
                a.Y[row * a.ldy + n] = fmaf(z * m, a.oscale, v);
              }
            }
          }
          continue;
        }
        if (EPI != 0) {
          // implicit-network epilogues (p2c_linear_act): both outputs leave through [32 rows x 32 channels] staging
          // tiles and TMA stores (rows >= M are clipped by the tensor maps)
          if (EPI == 1) {
            // H through the staging tile + TMA store; S = softplus'(Z) straight from registers (lanes = consecutive
            // channels of one row: a coalesced line per store instruction) - no second pass over the staging tile
            float* sp = a.s_tma ? a.S + (size_t)mrow * a.lds + n : nullptr;
            const float bl2e = a.beta * 1.4426950408889634f, inv_b = 0.6931471805599453f / a.beta;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float z = __uint_as_float(raw[j]) + bias;
              float h = z, sg = 1.f;                  // nn.Softplus(beta, threshold 20): identity above the threshold
              if (!(a.beta * z > 20.f) && !(dbg_mode & 16)) {
                // e = exp(beta z) <= e^20; softplus = ln(1 + e) / beta, sigmoid = e / (1 + e): MUFU ex2 / lg2 / rcp
                // (absolute error of h < 1e-9, relative error of sg < 1e-6 - below the 3xTF32 error of z itself)
                const float e = exp2f(bl2e * z);
                const float p = 1.f + e;
                h = (e < 1e-4f) ? (e - 0.5f * e * e) * (1.f / a.beta) : __log2f(p) * inv_b;
                sg = __fdividef(e, p);
              }
              st[j * 32 + lane] = h * a.oscale;
              if (sp && j < jmax && !(dbg_mode & 8)) sp[(size_t)j * a.lds] = sg;
            }
            fence_proxy_async();
            __syncwarp();
            if (!(dbg_mode & 32) && elect_one_sync()) {
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(reinterpret_cast<uint64_t>(&tmY)), "r"(smem_u32(st)), "r"(n0 + q * 32), "r"(mrow) : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          } else {
            const float* mp = a.mul + (size_t)mrow * a.ldmul + n;
            const float* vp = EPI >= 3 ? a.mul2 + (size_t)mrow * a.ldmul2 + n : nullptr;
            float* zp = EPI == 3 ? a.S + (size_t)mrow * a.lds + n : nullptr;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const bool ok = n_ok && j < jmax;
              const float m = ok ? __ldg(mp + (size_t)j * a.ldmul) : 0.f;   // warp = 128 B of row mrow+j
              const float z = __uint_as_float(raw[j]) + bias;
              if (EPI == 2) {
                st[j * 32 + lane] = z * m * a.oscale;
              } else {
                const float v = ok ? __ldg(vp + (size_t)j * a.ldmul2) : 0.f;
                if (EPI == 3) {
                  st[j * 32 + lane] = z * m * a.oscale;
                  if (ok) zp[(size_t)j * a.lds] = a.beta * z * v * (1.f - m);
                } else {
                  st[j * 32 + lane] = fmaf(z * m, a.oscale, v);
                }
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (elect_one_sync()) {
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(reinterpret_cast<uint64_t>(&tmY)), "r"(smem_u32(st)), "r"(n0 + q * 32), "r"(mrow) : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
          continue;
        }
        float* yp = (a.Y && !y_tma && n_ok) ? a.Y + (size_t)mrow * a.ldy + n : nullptr;
        float mx = NEG_INF, mn = POS_INF;
        // per-row-group bias (p2c_linear_group_bias): the 32 rows of a chunk lie in one group (bias_group % 32 == 0)
        const float bias_c = a.bias_rows ? (n_ok ? __ldg(a.bias_rows + (size_t)(mrow / a.bias_group) * a.N + n) : 0.f) : bias;
        if (jmax == 32) epi_chunk<true>(raw, bias_c, 32, y_tma ? st + lane : nullptr, yp, a.ldy, a.stats != nullptr, G != 0, t1, t2, mx, mn);
        else epi_chunk<false>(raw, bias_c, jmax, y_tma ? st + lane : nullptr, yp, a.ldy, a.stats != nullptr, G != 0, t1, t2, mx, mn);
        if (y_tma) {
          fence_proxy_async();
          __syncwarp();
          if (elect_one_sync()) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&tmY)), "r"(smem_u32(st)), "r"(n0 + q * 32), "r"(mrow) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (G) {
          gmx = fmaxf(gmx, mx); gmn = fminf(gmn, mn);
          const int rows_done = c * 32 + 32;
          if ((rows_done & (G - 1)) == 0) {            // G is 32, 64 or 128
            if (n_ok) {
              const size_t o = (size_t)((mrow + 32 - G) >> gshift) * a.N + n;
              a.Ymax[o] = gmx;
              a.Ymin[o] = gmn;
            }
            gmx = NEG_INF; gmn = POS_INF;
          }
        }
      }
      if (a.stats && n_ok) {
        atomicAdd(a.stats + n, (double)t1);
        atomicAdd(a.stats + a.N + n, (double)t2);
      }
    }
    __syncwarp();
    if (y_tma_all && elect_one_sync()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ W, int N, int K, float* __restrict__ out, int64_t ldw) {
  const int64_t total = (int64_t)N * ldw;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = e / ldw;
    const int k = (int)(e - n * ldw);
    const float w = k < K ? __ldg(W + n * K + k) : 0.f;
    const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
    out[e] = hi;
    out[total + e] = w - hi;
  }
}

// All streamed-weight layers of a forward in ONE launch: blockIdx.y = matrix (descriptors passed by value)
constexpr int SPLIT_MULTI_MAX = 16;
struct SplitMulti {
  const float* W[SPLIT_MULTI_MAX];
  float* out[SPLIT_MULTI_MAX];
  int N[SPLIT_MULTI_MAX], K[SPLIT_MULTI_MAX], ldw[SPLIT_MULTI_MAX];
  int sn[SPLIT_MULTI_MAX], sk[SPLIT_MULTI_MAX];   // element strides of W: (K, 1) as stored, (1, N) = its transpose
};
__global__ void __launch_bounds__(256)
split_tf32_multi_kernel(const SplitMulti d) {
  const int m = blockIdx.y;
  const float* __restrict__ W = d.W[m];
  float* __restrict__ out = d.out[m];
  const int K = d.K[m];
  const int64_t sn = d.sn[m], sk = d.sk[m];
  const int64_t ldw = d.ldw[m], total = (int64_t)d.N[m] * ldw;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = e / ldw;
    const int k = (int)(e - n * ldw);
    const float w = k < K ? __ldg(W + n * sn + k * sk) : 0.f;
    const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
    out[e] = hi;
    out[total + e] = w - hi;
  }
}

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ W, int N, int K, uint16_t* __restrict__ out, int64_t ldw) {
  const int64_t total = (int64_t)N * ldw;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = e / ldw;
    const int k = (int)(e - n * ldw);
    const float w = k < K ? __ldg(W + n * K + k) : 0.f;
    out[e] = (uint16_t)(pack_bf16(w, 0.f) & 0xffffu);
  }
}

int ss_stages(int KB, int y_tma, int* raw, int* xt, int raw_hi = 0) {
  const int tries_t[4][2] = {{4, 3}, {2, 3}, {4, 2}, {2, 2}};
  const int tries_r[4][2] = {{3, 3}, {3, 2}, {2, 2}, {2, 2}};   // raw_hi: a RAW stage lives as long as its XT / W stage
  const int (*tries)[2] = raw_hi ? tries_r : tries_t;
  for (int i = 0; i < 4; ++i)
    if (ss_smem_layout(KB, tries[i][0], tries[i][1], y_tma, raw_hi).total + 1024 <= 227 * 1024) {
      *raw = tries[i][0]; *xt = tries[i][1];
      return 1;
    }
  return 0;
}

}  // namespace

extern "C" int p2c_split_tf32(const float* W, int N, int K, float* out, int64_t ldw, void* stream) {
  if (!W || !out || N <= 0 || K <= 0 || ldw < K || (ldw % 4) != 0) return P2C_EINVAL;
  const int64_t total = (int64_t)N * ldw;
  const int blocks = (int)min((int64_t)148 * 8, (total + 255) / 256);
  split_tf32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, N, K, out, ldw);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_split_tf32_multi(const float* const* W, const int* N, const int* K, float* const* out,
                                    const int64_t* ldw, const int* transposed, const int64_t* src_ld, int count,
                                    void* stream) {
  if (!W || !N || !K || !out || !ldw || count <= 0) return P2C_EINVAL;
  if (count > SPLIT_MULTI_MAX) return P2C_EUNSUPPORTED;
  SplitMulti d;
  int64_t biggest = 0;
  for (int i = 0; i < count; ++i) {
    if (!W[i] || !out[i] || N[i] <= 0 || K[i] <= 0 || ldw[i] < K[i] || (ldw[i] % 4) != 0) return P2C_EINVAL;
    d.W[i] = W[i]; d.out[i] = out[i]; d.N[i] = N[i]; d.K[i] = K[i]; d.ldw[i] = (int)ldw[i];
    const bool tr = transposed && transposed[i];     // source stored as (K, N): the split of its transpose is written
    const int ld = src_ld ? (int)src_ld[i] : (tr ? N[i] : K[i]);   // row stride of the source as stored
    if (ld < (tr ? N[i] : K[i])) return P2C_EINVAL;
    d.sn[i] = tr ? 1 : ld;
    d.sk[i] = tr ? ld : 1;
    biggest = max(biggest, (int64_t)N[i] * ldw[i]);
  }
  for (int i = count; i < SPLIT_MULTI_MAX; ++i) { d.W[i] = nullptr; d.out[i] = nullptr; d.N[i] = d.K[i] = d.ldw[i] = d.sn[i] = d.sk[i] = 0; }
  dim3 grid((unsigned)min((int64_t)296, (biggest + 255) / 256), (unsigned)count);
  split_tf32_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_cast_bf16(const float* W, int N, int K, void* out, int64_t ldw, void* stream) {
  if (!W || !out || N <= 0 || K <= 0 || ldw < K || (ldw % 8) != 0) return P2C_EINVAL;
  const int64_t total = (int64_t)N * ldw;
  const int blocks = (int)min((int64_t)148 * 8, (total + 255) / 256);
  cast_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, N, K, reinterpret_cast<uint16_t*>(out), ldw);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

// 1 when the streamed-W tensor-core kernel takes the shape (given a pre-split / bf16 weight copy)
int p2c_linear_tc_ss_plan(int64_t ldx, int x_aligned16, int K, int has_mask, int pool_group, int precision) {
  if ((precision != P2C_PREC_3XTF32 && precision != P2C_PREC_BF16) || has_mask) return 0;
  if ((ldx % 4) != 0 || !x_aligned16 || K < 16) return 0;
  if (pool_group && pool_group != 32 && pool_group != 64 && pool_group != 128) return 0;
  int raw, xt;
  return ss_stages((K + TC_BK - 1) / TC_BK, 1, &raw, &xt);
}

struct SsEpiHost { int op; float beta, oscale; float* S; int64_t lds; const float* mul; int64_t ldmul;
                   const float* bias_rows = nullptr; int bias_group = 0;
                   const float* mul2 = nullptr; int64_t ldmul2 = 0; };

static int linear_tc_ss_launch(const float* X, int64_t ldx, const float* w_split, int64_t ldws, const float* bias,
                               const float* in_scale, const float* in_shift, float* Y, int64_t ldy, int M, int N, int K,
                               double* stats, int pool_group, float* Ymax, float* Ymin, int bf16,
                               const p2c_bn_fold* in_bn, const SsEpiHost& epi, cudaStream_t st);

int p2c_linear_tc_ss(const float* X, int64_t ldx, const float* w_split, int64_t ldws, const float* bias,
                     const float* in_scale, const float* in_shift, float* Y, int64_t ldy, int M, int N, int K,
                     double* stats, int pool_group, float* Ymax, float* Ymin, int bf16, const p2c_bn_fold* in_bn,
                     cudaStream_t st) {
  const SsEpiHost none{0, 0.f, 1.f, nullptr, 0, nullptr, 0};
  return linear_tc_ss_launch(X, ldx, w_split, ldws, bias, in_scale, in_shift, Y, ldy, M, N, K, stats, pool_group, Ymax,
                             Ymin, bf16, in_bn, none, st);
}

static int linear_tc_ss_launch(const float* X, int64_t ldx, const float* w_split, int64_t ldws, const float* bias,
                               const float* in_scale, const float* in_shift, float* Y, int64_t ldy, int M, int N, int K,
                               double* stats, int pool_group, float* Ymax, float* Ymin, int bf16,
                               const p2c_bn_fold* in_bn, const SsEpiHost& epi, cudaStream_t st) {
  const int KB = (K + TC_BK - 1) / TC_BK;
  const int y_tma = (Y && (ldy % 4) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0) ? 1 : 0;
  const int s_tma = (epi.op == 1 && epi.S) ? 1 : 0;
  if (epi.op != 0 && !y_tma) return P2C_EALIGN;
  if (s_tma && ((epi.lds % 4) != 0 || (reinterpret_cast<uintptr_t>(epi.S) & 15) != 0)) return P2C_EALIGN;
  int raw, xt;
  const int raw_hi = (!bf16 && !in_scale && !in_bn) ? 1 : 0;   // nothing to fold into the operand: RAW = hi operand
  if (!ss_stages(KB, y_tma, &raw, &xt, raw_hi)) return P2C_EUNSUPPORTED;
  if ((ldws % (bf16 ? 8 : 4)) != 0 || (reinterpret_cast<uintptr_t>(w_split) & 15) != 0) return P2C_EALIGN;
  CUtensorMap tmX, tmWhi, tmWlo, tmY;
  int rc;
  if ((rc = make_map_2d(&tmX, X, K, M, ldx, TC_BK, TC_BM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if (bf16) {
    // bf16 weight copy (N, ldws) bf16: boxes of [128 channels x 32 bf16 = 64 B], SWIZZLE_64B
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
    const cuuint64_t gstride[1] = {(cuuint64_t)ldws * 2};
    const cuuint32_t box[2] = {TC_BK, TC_BN};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&tmWhi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<float*>(w_split), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return (int)cudaErrorInvalidValue;
    tmWlo = tmWhi;
  } else {
    if ((rc = make_map_2d(&tmWhi, w_split, K, N, ldws, TC_BK, TC_BN, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_map_2d(&tmWlo, w_split + (size_t)N * ldws, K, N, ldws, TC_BK, TC_BN, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  tmY = tmX;
  if (y_tma && (rc = make_map_2d(&tmY, Y, N, M, ldy, 32, 32, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  CUtensorMap tmS = tmY;
  if (s_tma && (rc = make_map_2d(&tmS, epi.S, N, M, epi.lds, 32, 32, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  SsArgs a{bias, in_scale, in_shift, Y, ldy, M, N, K, KB, stats, pool_group, Ymax, Ymin, raw, xt,
           (M + TC_BM - 1) / TC_BM, (N + TC_BN - 1) / TC_BN, y_tma, p2c_bn_fold_dev(in_bn),
           epi.beta, epi.oscale, s_tma, epi.mul, epi.ldmul, epi.S, epi.lds, epi.mul2, epi.ldmul2, raw_hi, epi.bias_rows, epi.bias_group,
           getenv("P2C_SS_DBG") ? atoi(getenv("P2C_SS_DBG")) : 0};
  const SsSmem L = ss_smem_layout(KB, raw, xt, y_tma, raw_hi);
  int dev = 0;
  cudaGetDevice(&dev);
  static int sms_of[64] = {0};
  if (dev < 64 && sms_of[dev] == 0) {
    P2C_CUDA_TRY(cudaFuncSetAttribute(linear_tc_ss_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    P2C_CUDA_TRY(cudaFuncSetAttribute(linear_tc_ss_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    P2C_CUDA_TRY(cudaFuncSetAttribute(linear_tc_ss_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    P2C_CUDA_TRY(cudaFuncSetAttribute(linear_tc_ss_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    P2C_CUDA_TRY(cudaFuncSetAttribute(linear_tc_ss_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    P2C_CUDA_TRY(cudaFuncSetAttribute(linear_tc_ss_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms_of[dev] = n;
  }
  const int sms = p2c_sm_budget(dev < 64 ? sms_of[dev] : 148);
  const int tiles = a.m_tiles * a.n_tiles;
  const int grid = tiles < sms ? tiles : sms;
  const dim3 g3(grid), b3(SS_THREADS);
  const size_t sm = L.total + 1024;
  if (bf16) P2C_CUDA_TRY(p2c_launch(linear_tc_ss_kernel<true, 0>, g3, b3, sm, st, tmX, tmWhi, tmWlo, tmY, tmS, a));
  else if (epi.op == 1) P2C_CUDA_TRY(p2c_launch(linear_tc_ss_kernel<false, 1>, g3, b3, sm, st, tmX, tmWhi, tmWlo, tmY, tmS, a));
  else if (epi.op == 2) P2C_CUDA_TRY(p2c_launch(linear_tc_ss_kernel<false, 2>, g3, b3, sm, st, tmX, tmWhi, tmWlo, tmY, tmS, a));
  else if (epi.op == 3) P2C_CUDA_TRY(p2c_launch(linear_tc_ss_kernel<false, 3>, g3, b3, sm, st, tmX, tmWhi, tmWlo, tmY, tmS, a));
  else if (epi.op == 4) P2C_CUDA_TRY(p2c_launch(linear_tc_ss_kernel<false, 4>, g3, b3, sm, st, tmX, tmWhi, tmWlo, tmY, tmS, a));
  else P2C_CUDA_TRY(p2c_launch(linear_tc_ss_kernel<false, 0>, g3, b3, sm, st, tmX, tmWhi, tmWlo, tmY, tmS, a));
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

// p2c_linear with a bias per GROUP of rows: Y[m, :] = f(X[m, :]) W^T + bias_rows[m / group, :] (see include/point2cyl.h)
extern "C" int p2c_linear_group_bias(const float* X, int64_t ldx, const float* w_split, int64_t ldws,
                                     const float* bias_rows, int group, const float* in_scale, const float* in_shift,
                                     const p2c_bn_fold* in_bn, float* Y, int64_t ldy, int M, int N, int K, double* stats,
                                     void* stream) {
  if (!X || !w_split || !bias_rows || !Y || M <= 0 || N <= 0 || K <= 0 || ldx < K || ldy < N || ldws < K) return P2C_EINVAL;
  if (group <= 0 || group % 32 != 0) return P2C_EUNSUPPORTED;
  if ((in_scale == nullptr) != (in_shift == nullptr)) return P2C_EINVAL;
  if (int e = p2c_bn_fold_check(in_bn, K)) return e;
  if (!p2c_linear_tc_ss_plan(ldx, (reinterpret_cast<uintptr_t>(X) & 15) == 0, K, 0, 0, P2C_PREC_3XTF32))
    return P2C_EUNSUPPORTED;
  SsEpiHost epi{0, 0.f, 1.f, nullptr, 0, nullptr, 0};
  epi.bias_rows = bias_rows;
  epi.bias_group = group;
  return linear_tc_ss_launch(X, ldx, w_split, ldws, nullptr, in_scale, in_shift, Y, ldy, M, N, K, stats, 0, nullptr,
                             nullptr, 0, in_bn, epi, (cudaStream_t)stream);
}

// One hidden layer of the implicit sketch network on the tensor cores (3xTF32): see include/point2cyl.h
extern "C" int p2c_linear_act(const float* X, int64_t ldx, const float* w_split, int64_t ldws, const float* bias, int M,
                              int N, int K, int op, float beta, float oscale, float* Y, int64_t ldy, float* S,
                              int64_t lds, const float* Mul, int64_t ldmul, void* stream) {
  if (!X || !w_split || !Y || M <= 0 || N <= 0 || K <= 0 || ldx < K || ldy < N || ldws < K) return P2C_EINVAL;
  if (op < 0 || op > 2 || (op == 2 && (!Mul || ldmul < N)) || (S && (op != 1 || lds < N))) return P2C_EINVAL;
  if (!p2c_linear_tc_ss_plan(ldx, (reinterpret_cast<uintptr_t>(X) & 15) == 0, K, 0, 0, P2C_PREC_3XTF32))
    return P2C_EUNSUPPORTED;
  const SsEpiHost epi{op, beta, oscale, S, lds, Mul, ldmul};
  if (op == 0) {   // plain scaled layer: the multiplier epilogue needs a matrix, so scale through EPI 1? no: use EPI 0
    if (oscale != 1.f) return P2C_EUNSUPPORTED;
  }
  return linear_tc_ss_launch(X, ldx, w_split, ldws, bias, nullptr, nullptr, Y, ldy, M, N, K, nullptr, 0, nullptr, nullptr,
                             0, nullptr, epi, (cudaStream_t)stream);
}

// One layer of the implicit network's second-order backward on the tensor cores (3xTF32): see include/point2cyl.h
extern "C" int p2c_linear_act_bwd(const float* X, int64_t ldx, const float* w_split, int64_t ldws, int M, int N, int K,
                                  int op, float beta, float oscale, float* Y, int64_t ldy, const float* Mul,
                                  int64_t ldmul, const float* V, int64_t ldv, float* Z, int64_t ldz, void* stream) {
  if (!X || !w_split || !Y || !Mul || !V || M <= 0 || N <= 0 || K <= 0 || ldx < K || ldy < N || ldws < K) return P2C_EINVAL;
  if ((op != 3 && op != 4) || ldmul < N || ldv < N || (op == 3 && (!Z || ldz < N))) return P2C_EINVAL;
  if (!p2c_linear_tc_ss_plan(ldx, (reinterpret_cast<uintptr_t>(X) & 15) == 0, K, 0, 0, P2C_PREC_3XTF32))
    return P2C_EUNSUPPORTED;
  SsEpiHost epi{op, beta, oscale, op == 3 ? Z : nullptr, ldz, Mul, ldmul};
  epi.mul2 = V;
  epi.ldmul2 = ldv;
  return linear_tc_ss_launch(X, ldx, w_split, ldws, nullptr, nullptr, nullptr, Y, ldy, M, N, K, nullptr, 0, nullptr, nullptr,
                             0, nullptr, epi, (cudaStream_t)stream);
}
