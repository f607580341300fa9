// Per-point MLP layer on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   Y[m, n] = sum_k f(X[m, k]) * W[n, k] + bias[n]      f = identity | max(x*scale[k]+shift[k], 0)
//
// P2C_PREC_3XTF32 (fp32-faithful): every operand is split a = a_hi + a_lo with a_hi exactly
// representable in tf32 (low 13 mantissa bits cleared) and a_lo = a - a_hi (exact in fp32); the product is
// accumulated as a_hi*w_hi + a_lo*w_hi + a_hi*w_lo in fp32 TMEM accumulators (the dropped a_lo*w_lo
// term is ~2^-20 relative).  Three kind::tf32 MMAs per k-step.
//
// Persistent, warp-specialised CTA (384 threads, 1 CTA/SM), one n-tile of BN output channels per CTA:
//   warp 0      TMA producer: raw fp32 activation k-blocks [128 rows x 32 floats] -> smem ring
//               (cp.async.bulk.tensor.2d, SWIZZLE_128B, mbarrier complete_tx)
//   warps 8-11  operand transform (thread = row): smem -> registers, folded BatchNorm+ReLU of the
//               previous layer, hi/lo split, tcgen05.st into the A-operand TMEM ring
//   warp 1      MMA issuer (one elected lane): tcgen05.mma.cta_group::1.kind::tf32 with A from TMEM and
//               W_hi / W_lo from shared memory (K-major, SWIZZLE_128B, resident for the whole kernel);
//               tcgen05.commit releases A stages and publishes accumulators
//   warps 4-7   epilogue: tcgen05.ld accumulator -> +bias -> store Y, per-channel sum / sum-of-squares
//               (lane-transposing butterfly, 31 shuffles per 32 columns) and nsample max/min pool
//   warp 2      TMEM allocator
// Accumulators are double buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// TMEM map (columns): [0, 2*BN) accumulators, [2*BN, 2*BN + A_STAGES*64) A operand (hi 32 | lo 32).
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;            // fp32 elements per k-block = 128 bytes = one SWIZZLE_128B row
constexpr int TC_THREADS = 384;
constexpr int A_STAGES = 2;          // TMEM A-operand ring
constexpr int RAW_BYTES = TC_BM * TC_BK * 4;

struct TcArgs {
  const float* W; const float* bias;
  const float* in_scale; const float* in_shift;
  float* Y; int64_t ldy;
  int M, N, K, KB;
  double* stats;
  int pool_group;
  float* Ymax; float* Ymin;
  int raw_stages;
  int m_tiles;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
      "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
      "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B operand descriptor: 8-row atoms of 1024 B (SBO), 16-byte units, version 1 (sm_100).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                  // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                  // descriptor version
  d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

// Lane-transposing reduction: in  v[j] = this lane's (row's) value of column j,
//                             out v[0] = op over the 32 lanes of column `lane`.
template <class Op>
__device__ __forceinline__ void transpose_reduce(float (&v)[32], int lane, Op op) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool up = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half];
      const float keep = up ? v[i + half] : v[i];
      v[i] = op(keep, __shfl_xor_sync(P2C_FULL_MASK, send, half));
    }
  }
}
struct OpAdd { __device__ float operator()(float a, float b) const { return a + b; } };
struct OpMax { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct OpMin { __device__ float operator()(float a, float b) const { return fminf(a, b); } };

struct SmemLayout {
  uint32_t w_off, raw_off, scale_off, shift_off, pool_off, bar_off, total;
};
__host__ __device__ inline SmemLayout tc_smem_layout(int BN, int KB, int raw_stages) {
  SmemLayout L;
  uint32_t o = 0;
  L.w_off = o;      o += 2u * KB * BN * 128u;              // W_hi | W_lo, [KB][BN rows][128 B]
  L.raw_off = o;    o += (uint32_t)raw_stages * RAW_BYTES; // 1024-aligned: every term above is
  L.scale_off = o;  o += (uint32_t)KB * TC_BK * 4u;
  L.shift_off = o;  o += (uint32_t)KB * TC_BK * 4u;
  L.pool_off = o;   o += 4u * 2u * BN * 4u;                // [quarter][max|min][BN]
  L.bar_off = o;    o += 256u;
  L.total = o;
  return L;
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmA, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms (TMA destination, UMMA descriptors) need 1024-byte aligned shared addresses
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const SmemLayout L = tc_smem_layout(BN, a.KB, a.raw_stages);
  uint8_t* w_sm = smem + L.w_off;
  uint8_t* raw_sm = smem + L.raw_off;
  float* s_scale = reinterpret_cast<float*>(smem + L.scale_off);
  float* s_shift = reinterpret_cast<float*>(smem + L.shift_off);
  float* s_pool = reinterpret_cast<float*>(smem + L.pool_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* raw_full = bars;                 // [raw_stages] (<= 8)
  uint64_t* raw_empty = bars + 8;            // [raw_stages]
  uint64_t* a_full = bars + 16;              // [A_STAGES]
  uint64_t* a_empty = bars + 18;             // [A_STAGES]
  uint64_t* acc_full = bars + 20;            // [2]
  uint64_t* acc_empty = bars + 22;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * BN;
  const int KB = a.KB;
  const int RS = a.raw_stages;

  // ---- one-time setup -------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    for (int s = 0; s < RS; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 4); }
    for (int s = 0; s < A_STAGES; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  // weights: fp32 -> (hi, lo), K-major SWIZZLE_128B tiles, zero padded in n and k
  {
    const int kpad = KB * TC_BK;
    for (int e = tid; e < BN * kpad; e += TC_THREADS) {
      const int n = e / kpad, k = e - n * kpad;
      float w = 0.f;
      if (n0 + n < a.N && k < a.K) w = __ldg(a.W + (size_t)(n0 + n) * a.K + k);
      const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
      const float lo = w - hi;
      const int kb = k >> 5, kk = k & 31;
      const uint32_t off = (uint32_t)kb * BN * 128u + (uint32_t)n * 128u + ((((uint32_t)kk >> 2) ^ ((uint32_t)n & 7u)) << 4) +
                           ((uint32_t)kk & 3u) * 4u;
      *reinterpret_cast<float*>(w_sm + off) = hi;
      *reinterpret_cast<float*>(w_sm + (uint32_t)KB * BN * 128u + off) = lo;
    }
    for (int k = tid; k < kpad; k += TC_THREADS) {
      const bool ok = a.in_scale != nullptr && k < a.K;
      s_scale[k] = ok ? __ldg(a.in_scale + k) : (a.in_scale ? 0.f : 1.f);
      s_shift[k] = ok ? __ldg(a.in_shift + k) : 0.f;
    }
  }
  fence_proxy_async();   // generic-proxy smem writes (W tiles) -> visible to the tensor core's async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_acc = tmem_base;                       // + ab*BN
  const uint32_t tm_a = tmem_base + 2 * BN;                // + as*64 (+32 for lo)

  const int my_tiles = (a.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int m0 = ((int)blockIdx.x + t * (int)gridDim.x) * TC_BM;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&raw_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&raw_full[s], RAW_BYTES);
          tma_load_2d(raw_sm + (size_t)s * RAW_BYTES, &tmA, &raw_full[s], kb * TC_BK, m0);
          if (++s == RS) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BN);
      const uint32_t w_hi = smem_u32(w_sm);
      const uint32_t w_lo = w_hi + (uint32_t)KB * BN * 128u;
      int as = 0; uint32_t aph = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int ab = t & 1;
        const uint32_t accph = (uint32_t)(t >> 1) & 1u;
        mbar_wait(&acc_empty[ab], accph ^ 1);
        tc_fence_after();
        const uint32_t d = tm_acc + (uint32_t)ab * BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t ahi = tm_a + (uint32_t)as * 64u, alo = ahi + 32u;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bhi = make_kmajor_sw128_desc(w_hi + (uint32_t)kb * BN * 128u + ks * 32u);
            const uint64_t blo = make_kmajor_sw128_desc(w_lo + (uint32_t)kb * BN * 128u + ks * 32u);
            umma_tf32_ts(d, ahi + ks * 8u, bhi, idesc, (kb | ks) != 0);
            umma_tf32_ts(d, alo + ks * 8u, bhi, idesc, 1u);
            umma_tf32_ts(d, ahi + ks * 8u, blo, idesc, 1u);
          }
          umma_commit(&a_empty[as]);                 // frees this A stage once the MMAs above retire
          if (kb == KB - 1) umma_commit(&acc_full[ab]);
          if (++as == A_STAGES) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp >= 8) {
    // ===== operand transform: raw smem -> BN+ReLU -> hi/lo -> TMEM =====
    const int q = warp & 3;
    const int r = q * 32 + lane;                    // tile row == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const bool has_affine = a.in_scale != nullptr;
    int s = 0; uint32_t ph = 0;
    int as = 0; uint32_t aph = 0;
    for (int t = 0; t < my_tiles; ++t) {
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&raw_full[s], ph);
        const uint8_t* rowp = raw_sm + (size_t)s * RAW_BYTES + (size_t)r * 128;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 x = *reinterpret_cast<const float4*>(rowp + ((c ^ (r & 7)) << 4));
          float v[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float y = v[i];
            if (has_affine) {
              const int k = kb * TC_BK + c * 4 + i;
              y = fmaxf(fmaf(y, s_scale[k], s_shift[k]), 0.f);
            }
            const uint32_t h = __float_as_uint(y) & 0xffffe000u;
            hi[c * 4 + i] = h;
            lo[c * 4 + i] = __float_as_uint(y - __uint_as_float(h));
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[s]);   // raw slot consumed (values are in registers)
        if (++s == RS) { s = 0; ph ^= 1; }
        mbar_wait(&a_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t ta = tm_a + (uint32_t)as * 64u + lane_addr;
        tmem_st32(ta, hi);
        tmem_st32(ta + 32u, lo);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[as]);
        if (++as == A_STAGES) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int q = warp & 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    constexpr int NCH = BN / 32;
    double s1[NCH], s2[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) { s1[c] = 0.0; s2[c] = 0.0; }
    const float NEG_INF = -__int_as_float(0x7f800000), POS_INF = __int_as_float(0x7f800000);
    const int G = a.pool_group;
    for (int t = 0; t < my_tiles; ++t) {
      const int ab = t & 1;
      const uint32_t accph = (uint32_t)(t >> 1) & 1u;
      const int m0 = ((int)blockIdx.x + t * (int)gridDim.x) * TC_BM;
      const int row = m0 + q * 32 + lane;
      const bool valid = row < a.M;
      mbar_wait(&acc_full[ab], accph);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t raw[32];
        tmem_ld32(tm_acc + (uint32_t)ab * BN + (uint32_t)c * 32u + lane_addr, raw);
        tmem_wait_ld();
        float v[32];
        const int col0 = n0 + c * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float bj = (a.bias && col0 + j < a.N) ? __ldg(a.bias + col0 + j) : 0.f;
          v[j] = __uint_as_float(raw[j]) + bj;
        }
        if (a.Y && valid) {
          float* yr = a.Y + (size_t)row * a.ldy + col0;
          if (col0 + 32 <= a.N && ((reinterpret_cast<uintptr_t>(yr) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(yr + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < a.N) yr[j] = v[j];
          }
        }
        if (a.stats) {
          float tsum[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) tsum[j] = valid ? v[j] : 0.f;
          transpose_reduce(tsum, lane, OpAdd());
          s1[c] += (double)tsum[0];
#pragma unroll
          for (int j = 0; j < 32; ++j) tsum[j] = valid ? v[j] * v[j] : 0.f;
          transpose_reduce(tsum, lane, OpAdd());
          s2[c] += (double)tsum[0];
        }
        if (G) {
          float tm[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) tm[j] = valid ? v[j] : NEG_INF;
          transpose_reduce(tm, lane, OpMax());
          s_pool[(q * 2 + 0) * BN + c * 32 + lane] = tm[0];
#pragma unroll
          for (int j = 0; j < 32; ++j) tm[j] = valid ? v[j] : POS_INF;
          transpose_reduce(tm, lane, OpMin());
          s_pool[(q * 2 + 1) * BN + c * 32 + lane] = tm[0];
        }
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
      if (G) {
        asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps
        const int et = tid - 128;                         // 0..127
        const int wpg = G / 32;                           // warps (quarters) per pool group: 1, 2 or 4
        const int groups = 4 / wpg;
        for (int e = et; e < groups * BN; e += 128) {
          const int g = e / BN, col = e - g * BN;
          const int grow = m0 + g * G;
          if (grow < a.M && n0 + col < a.N) {
            float mx = NEG_INF, mn = POS_INF;
            for (int w = 0; w < wpg; ++w) {
              mx = fmaxf(mx, s_pool[((g * wpg + w) * 2 + 0) * BN + col]);
              mn = fminf(mn, s_pool[((g * wpg + w) * 2 + 1) * BN + col]);
            }
            const size_t o = (size_t)(grow / G) * a.N + n0 + col;
            a.Ymax[o] = mx;
            a.Ymin[o] = mn;
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // s_pool is rewritten by the next tile
      }
    }
    if (a.stats) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col = n0 + c * 32 + lane;
        if (col < a.N) {
          atomicAdd(a.stats + col, s1[c]);
          atomicAdd(a.stats + a.N + col, s2[c]);
        }
      }
    }
  }

  // ---- teardown ------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <int BN>
int launch_tc(const CUtensorMap& tm, TcArgs a, int n_tiles, cudaStream_t st) {
  const SmemLayout L = tc_smem_layout(BN, a.KB, a.raw_stages);
  auto k = linear_tc_kernel<BN>;
  P2C_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total + 1024));
  int sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int gx = sms / n_tiles;
  if (gx < 1) gx = 1;
  if (gx > a.m_tiles) gx = a.m_tiles;
  dim3 grid(gx, n_tiles);
  k<<<grid, TC_THREADS, L.total + 1024, st>>>(tm, a);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

}  // namespace

// Which kernel takes a given layer shape: 0 = fp32 SIMT, 1 = tcgen05 3xTF32.  Pure function of its arguments.
static int tc_plan(int64_t ldx, int x_aligned16, int M, int N, int K, int has_mask, int pool_group, int precision,
                   int* BN_out, int* KB_out, int* stages_out) {
  (void)M;
  if (precision != P2C_PREC_3XTF32) return 0;                      // bf16 variant: DESIGN.md, next
  if (has_mask) return 0;
  if ((ldx % 4) != 0 || !x_aligned16) return 0;                    // TMA global strides are multiples of 16 B
  if (K < 16) return 0;                                            // xyz-only first layers stay on the SIMT kernel
  if (pool_group && pool_group != 32 && pool_group != 64 && pool_group != 128) return 0;
  const int BN = N > 64 ? 128 : (N > 32 ? 64 : 32);
  const int KB = (K + TC_BK - 1) / TC_BK;
  int raw_stages = 4;
  while (raw_stages >= 2 && tc_smem_layout(BN, KB, raw_stages).total + 1024 > 227 * 1024) --raw_stages;
  if (raw_stages < 2) return 0;                                    // W not resident: SIMT (streaming-W variant next)
  if (BN_out) { *BN_out = BN; *KB_out = KB; *stages_out = raw_stages; }
  return 1;
}

extern "C" int p2c_linear_path(int64_t ldx, int M, int N, int K, int has_mask, int pool_group, int precision) {
  return tc_plan(ldx, 1, M, N, K, has_mask, pool_group, precision, nullptr, nullptr, nullptr);
}

int p2c_linear_tc(const float* X, int64_t ldx, const float* W, const float* bias, const float* in_scale,
                  const float* in_shift, const float* in_mask, int64_t ldmask, float* Y, int64_t ldy, int M, int N,
                  int K, double* stats, int pool_group, float* Ymax, float* Ymin, int precision, cudaStream_t st) {
  (void)ldmask;
  int BN, KB, raw_stages;
  if (!tc_plan(ldx, (reinterpret_cast<uintptr_t>(X) & 15) == 0, M, N, K, in_mask != nullptr, pool_group, precision,
               &BN, &KB, &raw_stages))
    return P2C_EUNSUPPORTED;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return (int)cudaErrorNotSupported;

  CUtensorMap tm;
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
  const cuuint64_t gstride[1] = {(cuuint64_t)ldx * 4};
  const cuuint32_t box[2] = {TC_BK, TC_BM};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;

  TcArgs a{W, bias, in_scale, in_shift, Y, ldy, M, N, K, KB, stats, pool_group, Ymax, Ymin, raw_stages,
           (M + TC_BM - 1) / TC_BM};
  const int n_tiles = (N + BN - 1) / BN;
  if (BN == 128) return launch_tc<128>(tm, a, n_tiles, st);
  if (BN == 64) return launch_tc<64>(tm, a, n_tiles, st);
  return launch_tc<32>(tm, a, n_tiles, st);
}
