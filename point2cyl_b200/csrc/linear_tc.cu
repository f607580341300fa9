// Per-point MLP layer on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   Y[m, n] = sum_k f(X[m, k]) * W[n, k] + bias[n]      f = identity | max(x*scale[k]+shift[k], 0)
//
// computed TRANSPOSED on the tensor core, D[n, m] = sum_k W[n, k] * f(X)[m, k]:
//   * the UMMA "A" operand is the weight matrix, resident in TENSOR MEMORY for the whole kernel
//     (lane = output channel n, column = k; hi | lo halves), so shared memory is free for a deep TMA ring;
//   * the UMMA "B" operand is the activation tile in shared memory, K-major SWIZZLE_128B - the layout a TMA
//     box load produces - after the operand transform (folded BatchNorm+ReLU of the previous layer, hi/lo split);
//   * the accumulator puts one OUTPUT CHANNEL per TMEM lane, i.e. per epilogue thread: bias, the BatchNorm
//     sum / sum-of-squares and the nsample max/min pool are plain in-register loops over the thread's row of
//     accumulator columns (no shuffles, no shared-memory staging), and Y[m, n0..n0+127] is one coalesced
//     512-byte row per store step.
//
// P2C_PREC_3XTF32 (fp32-faithful): a = a_hi + a_lo with a_hi exactly representable in tf32 (low 13 mantissa bits
// cleared) and a_lo = a - a_hi (exact); accumulated as w_hi*x_hi + w_lo*x_hi + w_hi*x_lo in fp32 (the dropped
// w_lo*x_lo term is ~2^-20 relative).  Three kind::tf32 MMAs (M=128, N=128, K=8) per k-step.
//
// Persistent warp-specialised CTA (384 threads, 1 CTA/SM); CTA (x, y) owns output channels [128y, 128y+128) and
// row tiles x, x+gridDim.x, ...:
//   warp 0      TMA producer: raw fp32 k-blocks [128 rows x 32 floats] -> RAW ring (cp.async.bulk.tensor.2d,
//               SWIZZLE_128B, mbarrier complete_tx)
//   warps 8-11  operand transform (thread = row): RAW ring -> registers -> BN+ReLU -> hi/lo -> XT ring
//               (same swizzled layout), fence.proxy.async, mbarrier arrive
//   warp 1      MMA issuer (one lane): tcgen05.mma.cta_group::1.kind::tf32, A = W from TMEM, B = XT from smem;
//               tcgen05.commit frees XT stages and publishes accumulators
//   warps 4-7   epilogue (thread = output channel): tcgen05.ld 32 accumulator columns (= 32 rows of Y) at a time;
//               Y leaves through per-warp [32 rows x 32 channels] staging tiles and TMA bulk tensor stores
//   warp 2      TMEM allocator;  warps 8-11 also load W into TMEM in the prologue
//
// TMEM columns: [0, 128*ACC) accumulators (ACC = 2 when K <= 128, else 1), then W_hi [KPAD] | W_lo [KPAD].
#include <cstdlib>
#include <cstring>

#include "bn_fold.cuh"
#include "tc_common.cuh"

using namespace p2c_tc;

namespace {

constexpr int LTC_THREADS = 512;   // 16 warps: producer, MMA, alloc, idle, 4 epilogue (even tiles), 4 transform, 4 epilogue (odd tiles)

struct TcArgs {
  const float* W; const float* bias;
  const float* in_scale; const float* in_shift;
  float* Y; int64_t ldy;
  int M, N, K, KB;
  double* stats;
  int pool_group;
  float* Ymax; float* Ymin;
  int raw_stages, xt_stages, acc_bufs;
  int m_tiles;
  int y_tma;        // Y rows are 16-byte aligned: per-warp TMA stores of [32 rows x 32 channels] boxes
  long long* dbg;   // optional timeline buffer (tools/tc_timeline.py); NULL in production
  int dbg_mode;     // tools only (env P2C_TC_DBG): bit0 skip transform math, bit1 skip epilogue body
  BnFoldDev bn;     // pending BatchNorm of X, folded in the prologue (bn.active) instead of in_scale / in_shift
  const int64_t* drop_seed;   // p = 0.5 dropout on f(X), drawn in the operand transform (the output heads, head.cu)
  P2cXyzFirst g;              // g.idx != NULL: the operand rows are recomputed from coordinates (common.cuh), X is not read
};

struct SmemLayout {
  uint32_t raw_off, xt_off, ystage_off, scale_off, shift_off, bits_off, w0_off, bar_off, total;
};
__host__ __device__ inline SmemLayout tc_smem_layout(int KB, int raw_stages, int xt_stages, int y_stage, int xyz_first = 0) {
  SmemLayout L;
  uint32_t o = 0;
  L.raw_off = o;    o += (uint32_t)raw_stages * RAW_BYTES;       // [stage][128 rows][128 B]
  L.xt_off = o;     o += (uint32_t)xt_stages * 2u * RAW_BYTES;   // [stage][hi|lo][128 rows][128 B]
  L.ystage_off = o; o += y_stage ? 8u * 2u * 4096u : 0u;         // [epilogue warp (2 groups x 4)][buf][32 rows][32 channels]
  L.scale_off = o;  o += (uint32_t)KB * TC_BK * 4u;
  L.shift_off = o;  o += (uint32_t)KB * TC_BK * 4u;
  L.bits_off = o;   o += 2u * TC_BM * 16u;                       // [buf][row] 128 dropout keep-bits (drop_seed) | centred xyz
  L.w0_off = o;     o += xyz_first ? (uint32_t)KB * TC_BK * 16u : 0u;   // [k] (W0[k][0..2], b0[k])
  L.bar_off = o;    o += 512u;
  L.total = o;
  return L;
}

// DBG = false (production): every timing toggle and timeline probe below is compiled out
template <bool DBG>
__global__ void __launch_bounds__(LTC_THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmY, const TcArgs a) {
  const int dbg_mode = DBG ? a.dbg_mode : 0;
  long long* const dbgp = DBG ? a.dbg : nullptr;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms (TMA destination, UMMA descriptors) need 1024-byte aligned shared addresses
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const bool xyz_first = a.g.idx != nullptr;
  const SmemLayout L = tc_smem_layout(a.KB, a.raw_stages, a.xt_stages, a.y_tma, xyz_first);
  uint8_t* raw_sm = smem + L.raw_off;
  uint8_t* xt_sm = smem + L.xt_off;
  uint8_t* ystage = smem + L.ystage_off;
  float* s_scale = reinterpret_cast<float*>(smem + L.scale_off);
  float* s_shift = reinterpret_cast<float*>(smem + L.shift_off);
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(smem + L.bits_off);
  float4* s_d = reinterpret_cast<float4*>(smem + L.bits_off);     // xyz-first: [buf][row] centred neighbour coordinates
  float4* s_w0 = reinterpret_cast<float4*>(smem + L.w0_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* raw_full = bars;                   // [MAX_RAW]
  uint64_t* raw_empty = bars + MAX_RAW;        // [MAX_RAW]
  uint64_t* xt_full = bars + 2 * MAX_RAW;      // [MAX_XT]
  uint64_t* xt_empty = xt_full + MAX_XT;       // [MAX_XT]
  uint64_t* acc_full = xt_empty + MAX_XT;      // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * TC_BN;
  const int KB = a.KB, KPAD = a.KB * TC_BK;
  const int RS = a.raw_stages, XS = a.xt_stages, ACC = a.acc_bufs;

  int dbg_k = 200;
  if (tid == 96) dbg_mark(dbgp, 3, dbg_k, 9000);          // kernel entry (idle warp 3)
  auto cta_mark = [&](int slot) {                          // per-CTA entry / prologue / roles-done times
    if (dbgp && tid == 96 && blockIdx.y == 0 && blockIdx.x < 256) {
      long long c;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(c));
      dbgp[2048 + blockIdx.x * 4 + slot] = c;
    }
  };
  cta_mark(0);

  // ---- one-time setup -------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    if (!xyz_first) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    for (int s = 0; s < RS; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 4); }
    for (int s = 0; s < XS; ++s) { mbar_init(&xt_full[s], 4); mbar_init(&xt_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  p2c_grid_dep_wait();          // everything above is local to the CTA; below: statistics / activations of predecessors
  p2c_grid_dep_launch();
  for (int k = tid; k < KPAD; k += LTC_THREADS) {
    float sc = 0.f, sh = 0.f;
    if (k < a.K) {
      const bool writer = blockIdx.x == 0 && blockIdx.y == 0;
      if (a.bn.active && xyz_first && a.g.moments && a.bn.stats) {
        double sum, sumsq;
        p2c_xyz_first_sums(a.g.moments, a.bn.count, (double)__ldg(a.g.W0 + (size_t)k * a.g.ldw0),
                           (double)__ldg(a.g.W0 + (size_t)k * a.g.ldw0 + 1), (double)__ldg(a.g.W0 + (size_t)k * a.g.ldw0 + 2),
                           a.g.b0 ? (double)__ldg(a.g.b0 + k) : 0.0, sum, sumsq);
        if (writer) {                         // leave the sums where the materialising kernel would have put them
          const_cast<double*>(a.bn.stats)[k] = sum;
          const_cast<double*>(a.bn.stats)[a.bn.C + k] = sumsq;
        }
        p2c_bn_fold_sums(a.bn, k, writer, sum, sumsq, sc, sh);
      } else if (a.bn.active) p2c_bn_fold_channel(a.bn, k, writer, sc, sh);
      else if (a.in_scale) { sc = __ldg(a.in_scale + k); sh = __ldg(a.in_shift + k); }
    }
    s_scale[k] = sc;
    s_shift[k] = sh;
    if (xyz_first) {
      // the first conv with its BatchNorm folded in: x = max(A . d + c, 0), A = scale w, c = scale b + shift (no
      // BatchNorm: A = w, c = b, no ReLU) - three FMAs per element in the operand transform
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < a.K) {
        w = make_float4(__ldg(a.g.W0 + (size_t)k * a.g.ldw0), __ldg(a.g.W0 + (size_t)k * a.g.ldw0 + 1),
                        __ldg(a.g.W0 + (size_t)k * a.g.ldw0 + 2), a.g.b0 ? __ldg(a.g.b0 + k) : 0.f);
        if (a.in_scale != nullptr || a.bn.active) w = make_float4(sc * w.x, sc * w.y, sc * w.z, fmaf(sc, w.w, sh));
      }
      // stored as the packed operands of the transform: per four channels 16 floats
      // [Ax0 Ax1 Ay0 Ay1 | Az0 Az1 c0 c1 | Ax2 Ax3 Ay2 Ay3 | Az2 Az3 c2 c3] - each LDS.128 yields two aligned pairs
      float* q = reinterpret_cast<float*>(s_w0) + (k >> 2) * 16 + ((k >> 1) & 1) * 8 + (k & 1);
      q[0] = w.x; q[2] = w.y; q[4] = w.z; q[6] = w.w;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_acc = tmem_base;                               // + ab*128
  const uint32_t tm_w = tmem_base + (uint32_t)ACC * TC_BM;         // W_hi at +k, W_lo at +KPAD+k

  // weights -> TMEM (thread = output channel; hi | lo), zero padded in n and k
  // N <= 64 channels in this CTA: lanes 64..127 hold a second copy of the weights, so the accumulator rows 64..127
  // repeat rows 0..63 (the M = 128 MMA costs the same either way) and the epilogue warps of lane quadrants 2 / 3 -
  // otherwise idle - drain the second half of each tile's rows: twice the epilogue throughput on the 64-wide layers
  const bool dup = (a.N - n0) <= 64 && (a.pool_group == 0 || a.pool_group == 32 || a.pool_group == 64);
  if (warp >= 8 && warp < 12) {
    const int q = warp & 3;
    const int n = n0 + (dup ? (q & 1) : q) * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const bool wvec = (a.K % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.W) & 15) == 0);
    for (int kb = 0; kb < KB; ++kb) {
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int k = kb * TC_BK + c * 4;
        float w[4] = {0.f, 0.f, 0.f, 0.f};
        if (n < a.N) {
          const float* wp = a.W + (size_t)n * a.K + k;
          if (wvec && k + 3 < a.K) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(wp));
            w[0] = t4.x; w[1] = t4.y; w[2] = t4.z; w[3] = t4.w;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (k + i < a.K) w[i] = __ldg(wp + i);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t h = __float_as_uint(w[i]) & 0xffffe000u;
          hi[c * 4 + i] = h;
          lo[c * 4 + i] = __float_as_uint(w[i] - __uint_as_float(h));
        }
      }
      tmem_st32(tm_w + lane_addr + (uint32_t)(kb * TC_BK), hi);
      tmem_st32(tm_w + lane_addr + (uint32_t)(KPAD + kb * TC_BK), lo);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 96) dbg_mark(dbgp, 3, dbg_k, 9001);          // prologue done
  cta_mark(1);

  const int my_tiles = (a.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    // ===== TMA producer (idle when the operand is recomputed from coordinates); warp-uniform loop, elected issuer =====
    if (!xyz_first) {
      int s = 0; uint32_t ph = 0;
      int dbg_n = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int m0 = ((int)blockIdx.x + t * (int)gridDim.x) * TC_BM;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&raw_empty[s], ph ^ 1);
          if (lane == 0) dbg_mark(dbgp, 3, dbg_n, t * 100 + kb);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&raw_full[s], RAW_BYTES);
            tma_load_2d(raw_sm + (size_t)s * RAW_BYTES, &tmA, &raw_full[s], kb * TC_BK, m0);
          }
          __syncwarp();
          if (++s == RS) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the loop (uniform control flow and operands), one elected lane issues =====
    {
      int xs = 0; uint32_t xph = 0;
      int dbg_n = 0;
      const bool mma_all = !(dbg_mode & (64 | 128));
      for (int t = 0; t < my_tiles; ++t) {
        const int ab = ACC == 2 ? (t & 1) : 0;
        const uint32_t accph = (ACC == 2 ? (uint32_t)(t >> 1) : (uint32_t)t) & 1u;
        mbar_wait(&acc_empty[ab], accph ^ 1);
        tc_fence_after();
        if (lane == 0) dbg_mark(dbgp, 0, dbg_n, t * 100 + 99);
        const uint32_t d = tm_acc + (uint32_t)ab * TC_BM;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&xt_full[xs], xph);
          tc_fence_after();
          if (lane == 0) dbg_mark(dbgp, 0, dbg_n, t * 100 + kb);
          const uint32_t x_hi = smem_u32(xt_sm + (size_t)xs * 2 * RAW_BYTES);
          // descriptors of the four 8-column k-steps differ in the start-address field only (+32 bytes = +2 units)
          const uint64_t bhi0 = make_kmajor_sw128_desc(x_hi), blo0 = make_kmajor_sw128_desc(x_hi + RAW_BYTES);
          const uint32_t w_hi = tm_w + (uint32_t)(kb * TC_BK), w_lo = w_hi + (uint32_t)KPAD;
          if (elect_one_sync()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma_tf32_ts(d, w_hi + ks * 8u, bhi0 + (uint64_t)(ks * 2), TC_IDESC, (kb | ks) != 0);
              if (mma_all) {
                umma_tf32_ts(d, w_lo + ks * 8u, bhi0 + (uint64_t)(ks * 2), TC_IDESC, 1u);
                umma_tf32_ts(d, w_hi + ks * 8u, blo0 + (uint64_t)(ks * 2), TC_IDESC, 1u);
              }
            }
            umma_commit(&xt_empty[xs]);                // frees this XT stage once the MMAs above retire
            if (kb == KB - 1) umma_commit(&acc_full[ab]);
          }
          __syncwarp();
          if (++xs == XS) { xs = 0; xph ^= 1; }
        }
      }
    }
  } else if (warp >= 8 && warp < 12) {
    // ===== operand transform: RAW ring -> BN+ReLU -> hi/lo -> XT ring =====
    // thread = (16-byte column chunk cj, row residue r7 = row mod 8, half hf): its eight rows are r7 + 8 (8 hf + i).
    // * its four k columns are fixed, so the folded BatchNorm scale / shift are 8 registers per k-block;
    // * row mod 8 is fixed, so the SWIZZLE_128B position of its chunk is ONE per-thread constant and the eight rows are
    //   1024 bytes apart: every shared-memory access below is [base + immediate] (no address arithmetic per row);
    // * a warp covers 4 full 128-byte rows per access: conflict-free.
    const int tt = tid - 256;
    const int cj = tt & 7, r7 = (tt >> 3) & 7, hf = tt >> 6;
    const uint32_t toff = (uint32_t)(hf * 8192 + r7 * 128 + ((cj ^ r7) << 4));
    const int row0 = hf * 64 + r7;                        // its rows: row0 + 8 i
    const bool has_affine = (a.in_scale != nullptr || a.bn.active) && !(dbg_mode & 1);
    int s = 0; uint32_t ph = 0;
    int xs = 0; uint32_t xph = 0;
    int dbg_n = 0;
    int bits_gen = 0;
    // hi = low 13 mantissa bits cleared, lo = x - hi (exact), stored at the thread's swizzled position of both tiles
    auto split_store = [&](uint8_t* hip, int i, float4 x) {
      float4 h;
      h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
      h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
      h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
      h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
      const float2 l0 = __fadd2_rn(make_float2(x.x, x.y), make_float2(-h.x, -h.y));
      const float2 l1 = __fadd2_rn(make_float2(x.z, x.w), make_float2(-h.z, -h.w));
      *reinterpret_cast<float4*>(hip + i * 1024) = h;
      *reinterpret_cast<float4*>(hip + RAW_BYTES + i * 1024) = make_float4(l0.x, l0.y, l1.x, l1.y);
    };
    auto affine_relu = [&](float4& x, const float4& sc, const float4& sh) {
      const float2 p0 = __ffma2_rn(make_float2(x.x, x.y), make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
      const float2 p1 = __ffma2_rn(make_float2(x.z, x.w), make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
      x = make_float4(fmaxf(p0.x, 0.f), fmaxf(p0.y, 0.f), fmaxf(p1.x, 0.f), fmaxf(p1.y, 0.f));
    };
    if (xyz_first) {
      // ----- operand recomputed from coordinates: thread tt owns row tt of the tile for the gather (index two tiles
      // ahead, coordinates one tile ahead: the dependent idx -> xyz chain never stalls the tile being produced), then
      // every thread computes its (8 rows x 4 channels) patch per k-block with the BatchNorm folded into the conv:
      // x = max(A_z d_z + (A_y d_y + (A_x d_x + c)), 0) - packed FFMA2 on channel pairs
      const P2cXyzFirst& g = a.g;
      auto tile_row = [&](int t) { return (int64_t)((int)blockIdx.x + t * (int)gridDim.x) * TC_BM + tt; };
      // volatile asm loads: issued where they are written (the compiler may not sink them down to their first use,
      // which would expose the whole idx -> xyz latency every tile)
      auto ldg_f = [](const float* p) { float v; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; };
      auto load_idx = [&](int t) -> int64_t {
        const int64_t r = tile_row(t);
        int64_t v = 0;
        if (t < my_tiles && r < a.M) asm volatile("ld.global.nc.s64 %0, [%1];" : "=l"(v) : "l"(g.idx + r));
        return v;
      };
      // the six loaded coordinates stay raw in registers until the NEXT tile subtracts them: the first instruction that
      // reads a load's result is a whole tile later (an in-order warp stalls at the first use, not at the load)
      struct Raw6 { float px, py, pz, cx, cy, cz; };
      auto load_pc = [&](int t, int64_t p) -> Raw6 {
        const int64_t r = tile_row(t);
        if (t >= my_tiles || r >= a.M || (dbg_mode & 16)) return Raw6{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const unsigned bs = (unsigned)r / (unsigned)g.ns, b = bs / (unsigned)g.S;
        p = (p < 0 || p >= g.N) ? 0 : p;
        const float* pp = g.xyz + ((size_t)b * g.N + (size_t)p) * 3;
        const float* cc = g.new_xyz + (size_t)bs * 3;
        return Raw6{ldg_f(pp), ldg_f(pp + 1), ldg_f(pp + 2), ldg_f(cc), ldg_f(cc + 1), ldg_f(cc + 2)};
      };
      int64_t i1 = load_idx(1);
      Raw6 cur = load_pc(0, load_idx(0));
      for (int t = 0; t < my_tiles; ++t) {
        const int64_t i2 = load_idx(t + 2);
        const Raw6 nxt = load_pc(t + 1, i1);
        float4* dbuf = s_d + (t & 1) * TC_BM;
        dbuf[tt] = make_float4(cur.px - cur.cx, cur.py - cur.cy, cur.pz - cur.cz, 0.f);
        asm volatile("bar.sync 2, 128;" ::: "memory");
        float4 d[8];                                   // the centred coordinates of this thread's eight rows: once per tile
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = dbuf[row0 + 8 * i];
        for (int kb = 0; kb < KB; ++kb) {
          // channel pairs as packed operands: (A_x, A_y, A_z, c) of channels 4 cj + {0,1} and + {2,3}
          const float4* wq = s_w0 + (kb * (TC_BK / 4) + cj) * 4;
          const float4 p0 = wq[0], p1 = wq[1], p2 = wq[2], p3 = wq[3];
          const float2 ax01 = make_float2(p0.x, p0.y), ay01 = make_float2(p0.z, p0.w), az01 = make_float2(p1.x, p1.y),
                       c01 = make_float2(p1.z, p1.w);
          const float2 ax23 = make_float2(p2.x, p2.y), ay23 = make_float2(p2.z, p2.w), az23 = make_float2(p3.x, p3.y),
                       c23 = make_float2(p3.z, p3.w);
          mbar_wait(&xt_empty[xs], xph ^ 1);
          uint8_t* hip = xt_sm + (size_t)xs * 2 * RAW_BYTES + toff;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 x;
            if (dbg_mode & 1) { x = d[i]; } else {
              const float2 dx = make_float2(d[i].x, d[i].x), dy = make_float2(d[i].y, d[i].y), dz = make_float2(d[i].z, d[i].z);
              const float2 y01 = __ffma2_rn(az01, dz, __ffma2_rn(ay01, dy, __ffma2_rn(ax01, dx, c01)));
              const float2 y23 = __ffma2_rn(az23, dz, __ffma2_rn(ay23, dy, __ffma2_rn(ax23, dx, c23)));
              x = has_affine ? make_float4(fmaxf(y01.x, 0.f), fmaxf(y01.y, 0.f), fmaxf(y23.x, 0.f), fmaxf(y23.y, 0.f))
                             : make_float4(y01.x, y01.y, y23.x, y23.y);
            }
            split_store(hip, i, x);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&xt_full[xs]);
          if (++xs == XS) { xs = 0; xph ^= 1; }
        }
        cur = nxt;
        i1 = i2;
      }
    } else
    for (int t = 0; t < my_tiles; ++t) {
      const int m0 = ((int)blockIdx.x + t * (int)gridDim.x) * TC_BM;
      for (int kb = 0; kb < KB; ++kb) {
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_affine) {
          sc = *reinterpret_cast<const float4*>(s_scale + kb * TC_BK + cj * 4);
          sh = *reinterpret_cast<const float4*>(s_shift + kb * TC_BK + cj * 4);
        }
        const uint32_t* bits = nullptr;
        if (a.drop_seed) {
          // dropout keep-bits of the tile's rows for channels [128 (kb / 4), +128): thread tt draws row tt's four
          // words (one Philox4x32-10 call, the counter of head.cu / p2c_head_bwd), the four transform warps meet on a
          // named barrier; two buffers, so the next draw never overwrites words a slower warp still reads
          if ((kb & 3) == 0) {
            const P2CPhilox4 b4 = p2c_dropout_bits(a.drop_seed, (int64_t)m0 + tt, kb >> 2);
            *reinterpret_cast<uint4*>(s_bits + ((bits_gen & 1) * TC_BM + tt) * 4) = make_uint4(b4.v[0], b4.v[1], b4.v[2], b4.v[3]);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            ++bits_gen;
          }
          bits = s_bits + ((bits_gen - 1) & 1) * TC_BM * 4 + (kb & 3);
        }
        mbar_wait(&raw_full[s], ph);
        if (tid == 256) dbg_mark(dbgp, 1, dbg_n, t * 100 + kb);
        const uint8_t* rawp = raw_sm + (size_t)s * RAW_BYTES + toff;
        float4 x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = *reinterpret_cast<const float4*>(rawp + i * 1024);
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[s]);   // raw slot consumed (values are in registers)
        if (++s == RS) { s = 0; ph ^= 1; }
        if (has_affine) {
#pragma unroll
          for (int i = 0; i < 8; ++i) affine_relu(x[i], sc, sh);
        }
        if (bits) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t w = bits[(row0 + 8 * i) * 4] >> (cj * 4);   // keep-bits of channels 32 (kb & 3) + 4 cj ..+3
            x[i].x = (w & 1u) ? 2.f * x[i].x : 0.f;
            x[i].y = (w & 2u) ? 2.f * x[i].y : 0.f;
            x[i].z = (w & 4u) ? 2.f * x[i].z : 0.f;
            x[i].w = (w & 8u) ? 2.f * x[i].w : 0.f;
          }
        }
        mbar_wait(&xt_empty[xs], xph ^ 1);
        uint8_t* hip = xt_sm + (size_t)xs * 2 * RAW_BYTES + toff;
        if (!(dbg_mode & 512)) {
#pragma unroll
          for (int i = 0; i < 8; ++i) split_store(hip, i, x[i]);
        }
        if (!(dbg_mode & 256))
        fence_proxy_async();                         // generic-proxy smem writes -> tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&xt_full[xs]);
        if (tid == 256) dbg_mark(dbgp, 1, dbg_n, t * 100 + kb + 50);
        if (++xs == XS) { xs = 0; xph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread = output channel n; accumulator columns = rows of Y =====
    // Two groups of four warps (warps 4-7 and 12-15; warp % 4 = TMEM lane quadrant): with two accumulator buffers
    // group g drains the tiles t = g (mod 2), so two tiles are in the epilogue at once.  One warp per scheduler
    // issues the ~260 dependent instructions of a 32-row chunk at ~0.27 IPC (ncu: 36 % fixed-latency waits), which
    // made the epilogue - not HBM, not the tensor pipe - pace the K = 64 layers; a second warp per scheduler
    // overlaps those stalls.
    const int grp = warp >= 12 ? 1 : 0;
    const int q = warp & 3;
    const int cq = dup ? (q & 1) : q;                  // 32-channel group of this warp
    const int c_lo = dup ? (q >> 1) * 2 : 0, c_hi = dup ? c_lo + 2 : 4;   // its 32-row chunks of a tile
    const int ch = cq * 32 + lane;                     // channel within the CTA's 128
    const int n = n0 + ch;
    const bool n_ok = n < a.N;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float bias = (a.bias && n_ok) ? __ldg(a.bias + n) : 0.f;
    const float NEG_INF = -__int_as_float(0x7f800000), POS_INF = __int_as_float(0x7f800000);
    const int G = a.pool_group;
    const int gshift = G ? 31 - __clz(G) : 0;
    float gmx = NEG_INF, gmn = POS_INF;                // running pool over the current group
    int dbg_n = 0;
    int ybuf = 0;
    float* ystg = reinterpret_cast<float*>(ystage + (size_t)(grp * 4 + q) * 8192);   // this warp's two 4 KB staging tiles
    // a [32 x 32] box sticking out over channel N is stored by the warp itself (TMA stores clip at 16-byte granularity)
    const bool y_tma = a.Y != nullptr && a.y_tma && (n0 + cq * 32 + 32 <= a.N);
    const bool warp_live = n0 + cq * 32 < a.N;         // warp-uniform
    // BatchNorm sums across tiles as float-float pairs (Knuth two-sum: the rounding error of every addition is
    // carried in the low word), converted to float64 once at the end - the fp64 pipe is slow here: two DADDs every
    // four tiles showed up as 5 % of the kernel's stall samples (math-pipe throttle)
    float h1 = 0.f, l1 = 0.f, h2 = 0.f, l2 = 0.f;
    auto two_sum = [](float& hi, float& lo, float t) {
      const float s_ = __fadd_rn(hi, t);
      const float bb = __fadd_rn(s_, -hi);
      lo = __fadd_rn(lo, __fadd_rn(__fadd_rn(hi, -__fadd_rn(s_, -bb)), __fadd_rn(t, -bb)));
      hi = s_;
    };
    for (int t = (ACC == 2 ? grp : 0); t < ((ACC == 2 || grp == 0) ? my_tiles : 0); t += (ACC == 2 ? 2 : 1)) {
      const int ab = ACC == 2 ? (t & 1) : 0;
      const uint32_t accph = (ACC == 2 ? (uint32_t)(t >> 1) : (uint32_t)t) & 1u;
      const int m0 = ((int)blockIdx.x + t * (int)gridDim.x) * TC_BM;
      mbar_wait(&acc_full[ab], accph);
      tc_fence_after();
      if (ch == 0 && grp == 0) dbg_mark(dbgp, 2, dbg_n, t * 100 + 99);
      float t1 = 0.f, t2 = 0.f;
      if (!warp_live) {
        // every channel of this warp lies beyond N (e.g. N = 64 on a 128-lane tile): nothing to read or reduce -
        // only hand the accumulator back, which also leaves the TMEM read port to the live warps
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[ab]);
        continue;
      }
      // The chunk loop is deliberately NOT unrolled: one chunk is ~250 straight-line instructions (4 KB); unrolled
      // four times (x the FULL / partial variants) the epilogue alone overflowed the instruction cache and the
      // whole kernel ran at ~0.25 IPC (K = 64 layers: 177 us -> see profiles/README.md).
      const uint32_t acc_addr = tm_acc + (uint32_t)ab * TC_BM + lane_addr;
#pragma unroll 1
      for (int c = c_lo; c < c_hi; ++c) {
        uint32_t raw[32];
        if (!(dbg_mode & 1024)) {
        tmem_ld32(acc_addr + (uint32_t)c * 32u, raw);
        tmem_wait_ld();
        }
        if (c == c_hi - 1) {                           // this warp's part is read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[ab]);
        }
        if (dbg_mode & 2) continue;
        const int mrow = m0 + c * 32;
        const int jmax = min(32, a.M - mrow);          // rows past M hold garbage (TMA zero fill + affine)
        if (jmax <= 0) continue;
        float* st = ystg + ybuf * 1024;
        if (y_tma) {
          // the bulk store that read this staging tile two chunks ago must have drained it
          if (elect_one_sync()) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // (the same lane every time)
          __syncwarp();
        }
        float* yp = (a.Y && !y_tma && n_ok) ? a.Y + (size_t)mrow * a.ldy + n : nullptr;
        float mx = NEG_INF, mn = POS_INF;
        if (jmax == 32) epi_chunk<true>(raw, bias, 32, y_tma ? st + lane : nullptr, yp, a.ldy, a.stats != nullptr && !(dbg_mode & 4), G != 0 && !(dbg_mode & 8), t1, t2, mx, mn);
        else epi_chunk<false>(raw, bias, jmax, y_tma ? st + lane : nullptr, yp, a.ldy, a.stats != nullptr, G != 0, t1, t2, mx, mn);
        if (y_tma) {
          // [32 rows][32 channels] staged (lanes = channels: conflict-free); the TMA engine writes it out:
          // one bulk tensor store per warp and chunk instead of 32 LSU store instructions
          fence_proxy_async();
          __syncwarp();
          if (elect_one_sync()) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&tmY)), "r"(smem_u32(st)), "r"(n0 + cq * 32), "r"(mrow) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ybuf ^= 1;
        }
        if (G) {
          // pool groups are 32, 64 or 128 consecutive rows and tiles start on group boundaries
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
        if (ch == 0 && grp == 0) dbg_mark(dbgp, 2, dbg_n, t * 100 + c);
      }
      two_sum(h1, l1, t1);
      two_sum(h2, l2, t2);
    }
    const double s1 = (double)h1 + (double)l1, s2 = (double)h2 + (double)l2;
    __syncwarp();
    if (y_tma && elect_one_sync()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (a.stats && n_ok) {
      atomicAdd(a.stats + n, s1);
      atomicAdd(a.stats + a.N + n, s2);
    }
  }

  // ---- teardown ------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (tid == 96) dbg_mark(dbgp, 3, dbg_k, 9002);          // all roles finished
  cta_mark(2);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

long long* g_tc_dbg = nullptr;

struct TcPlan { int KB, raw_stages, xt_stages, acc_bufs; };

inline int tc_raw_stages(int KB, int xt, int y_stage) {
  int raw = MAX_RAW;
  while (raw >= 2 && tc_smem_layout(KB, raw, xt, y_stage).total + 1024 > 227 * 1024) --raw;
  return raw;
}

// Which kernel takes a given layer shape: 0 = fp32 SIMT, 1 = tcgen05 3xTF32.  Pure function of its arguments.
int tc_plan(int64_t ldx, int x_aligned16, int M, int N, int K, int has_mask, int pool_group, int precision,
            TcPlan* out) {
  (void)M; (void)N;
  if (precision != P2C_PREC_3XTF32) return 0;                      // bf16: streamed-weight kernel (linear_tc_ss.cu)
  if (has_mask) return 0;
  if ((ldx % 4) != 0 || !x_aligned16) return 0;                    // TMA global strides are multiples of 16 B
  if (K < 16) return 0;                                            // xyz-only first layers stay on the SIMT kernel
  if (pool_group && pool_group != 32 && pool_group != 64 && pool_group != 128) return 0;
  const int KB = (K + TC_BK - 1) / TC_BK;
  const int KPAD = KB * TC_BK;
  int acc = 0;
  if (2 * KPAD + 2 * TC_BM <= 512) acc = 2;                        // TMEM: accumulators + W_hi + W_lo
  else if (2 * KPAD + TC_BM <= 512) acc = 1;
  else return 0;                                                   // K > 192: W does not fit TMEM (SIMT for now)
  const int xt = 3;
  const int raw = tc_raw_stages(KB, xt, 1);
  if (raw < 2) return 0;
  if (out) { out->KB = KB; out->raw_stages = raw; out->xt_stages = xt; out->acc_bufs = acc; }
  return 1;
}

}  // namespace

// tools only: timeline buffer (4 roles x 256 x {tag, globaltimer ns} + 256 x 4 per-CTA times)
extern "C" int p2c_debug_set_timeline(void* buf) { g_tc_dbg = reinterpret_cast<long long*>(buf); return 0; }

int p2c_linear_tc_ss_plan(int64_t ldx, int x_aligned16, int K, int has_mask, int pool_group, int precision);

extern "C" int p2c_linear_path(int64_t ldx, int M, int N, int K, int has_mask, int pool_group, int precision,
                               int has_split) {
  if (tc_plan(ldx, 1, M, N, K, has_mask, pool_group, precision, nullptr)) return 1;
  if (has_split && p2c_linear_tc_ss_plan(ldx, 1, K, has_mask, pool_group, precision))
    return precision == P2C_PREC_BF16 ? 3 : 2;
  return 0;
}

int p2c_linear_tc(const float* X, int64_t ldx, const float* W, const float* bias, const float* in_scale,
                  const float* in_shift, const float* in_mask, int64_t ldmask, float* Y, int64_t ldy, int M, int N,
                  int K, double* stats, int pool_group, float* Ymax, float* Ymin, int precision,
                  const p2c_bn_fold* in_bn, cudaStream_t st, const int64_t* drop_seed, const P2cXyzFirst* xyz_first) {
  (void)ldmask;
  TcPlan p;
  const bool gat = xyz_first != nullptr;     // the operand is recomputed from coordinates: X / ldx are not used
  if (!tc_plan(gat ? 4 : ldx, gat || (reinterpret_cast<uintptr_t>(X) & 15) == 0, M, N, K, in_mask != nullptr,
               pool_group, precision, &p))
    return P2C_EUNSUPPORTED;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return (int)cudaErrorNotSupported;

  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = CUDA_SUCCESS;
  if (!gat) {
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t gstride[1] = {(cuuint64_t)ldx * 4};
    const cuuint32_t box[2] = {TC_BK, TC_BM};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
  }

  CUtensorMap tmY = tm;
  int y_tma = 0;
  if (Y && (ldy % 4) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0) {
    // rows of Y beyond M and channels beyond N are clipped by the TMA unit
    const cuuint64_t ydim[2] = {(cuuint64_t)N, (cuuint64_t)M};
    const cuuint64_t ystride[1] = {(cuuint64_t)ldy * 4};
    const cuuint32_t ybox[2] = {32, 32};
    r = enc(&tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, Y, ydim, ystride, ybox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
    y_tma = 1;
  }
  const char* dm = getenv("P2C_TC_DBG");
  p.raw_stages = gat ? 2 : tc_raw_stages(p.KB, p.xt_stages, y_tma);
  TcArgs a{W, bias, in_scale, in_shift, Y, ldy, M, N, K, p.KB, stats, pool_group, Ymax, Ymin,
           p.raw_stages, p.xt_stages, p.acc_bufs, (M + TC_BM - 1) / TC_BM, y_tma, g_tc_dbg, dm ? atoi(dm) : 0,
           p2c_bn_fold_dev(in_bn), drop_seed, gat ? *xyz_first : P2cXyzFirst{}};
  const SmemLayout L = tc_smem_layout(p.KB, p.raw_stages, p.xt_stages, y_tma, gat);
  int dev = 0;
  cudaGetDevice(&dev);
  static int sms_of[64] = {0};   // per-device one-time setup (idempotent, so a race is harmless)
  if (dev < 64 && sms_of[dev] == 0) {
    P2C_CUDA_TRY(cudaFuncSetAttribute(linear_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    P2C_CUDA_TRY(cudaFuncSetAttribute(linear_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms_of[dev] = n;
  }
  const int sms = p2c_sm_budget(dev < 64 ? sms_of[dev] : 148);
  const int n_tiles = (N + TC_BN - 1) / TC_BN;
  int gx = sms / n_tiles;
  if (gx < 1) gx = 1;
  if (gx > a.m_tiles) gx = a.m_tiles;
  dim3 grid(gx, n_tiles);
  if (a.dbg || a.dbg_mode) P2C_CUDA_TRY(p2c_launch(linear_tc_kernel<true>, grid, dim3(LTC_THREADS), L.total + 1024, st, tm, tmY, a));
  else P2C_CUDA_TRY(p2c_launch(linear_tc_kernel<false>, grid, dim3(LTC_THREADS), L.total + 1024, st, tm, tmY, a));
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
