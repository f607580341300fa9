// A whole set-abstraction level as ONE tcgen05 kernel (north_star: "the per-point MLP stack as one fused kernel"):
//
//   gather  ->  conv0 (3 -> 64) + BN + ReLU  ->  conv1 (64 -> 64) + BN + ReLU  ->  conv2 (64 -> C2 <= 128) + BN + ReLU
//           ->  max over the nsample neighbours                       models/pointnet_util.py:130-139, 181-207
//
// for a level WITHOUT input features (sa1 of the backbone) whose BatchNorm layers use statistics that are known before
// the launch - eval mode (running statistics: eval.py:215,231,268), or batch statistics supplied by the caller.  Only
// the ball-query indices and the coordinates are read and only the pooled (B*S, C2) rows are written: no activation of
// the level ever reaches HBM (the layer-by-layer path writes and re-reads (B*S*ns, 64) fp32 = 2 x 268 MB at config 2,
// and launches three kernels).  Train-mode BatchNorm needs the batch statistics of conv1's output before conv2 can be
// applied - a grid-wide dependency - so training keeps one kernel per layer (linear_tc.cu).
//
// Per tile of 128 rows (= 128 / ns pool groups), persistent CTA, 512 threads, warp-specialised:
//   warps 8-11   L1: thread tt gathers row tt's index / coordinates two / one tile ahead, then every thread evaluates
//                relu(A d + c) (conv0 with BN folded in, packed FMAs) for its (8 rows x 4 channels) of both k-blocks,
//                splits into tf32 hi | lo and stores the K-major SWIZZLE_128B tiles of the XT1 ring (as linear_tc.cu)
//   warp 1       MMA issuer, software-pipelined  L2(t+1), L3(t):
//                L2: D2[row, ch] = X1[row, :] W1[ch, :]^T - ROWS are the UMMA M (128 TMEM lanes), the 64 channels its N:
//                    both operands from shared memory (W1 hi | lo resident for the whole kernel, pre-swizzled); half the
//                    tensor time of the channels-as-lanes form, which would pad 64 channels to 128 lanes
//                L3: D3[ch, row] = W2[ch, :] X2[row, :]^T - channels are the lanes (A = W2 hi | lo resident in TENSOR
//                    MEMORY), so the pool over rows is an in-register loop per epilogue thread
//                3xTF32 both (hi*hi + lo*hi + hi*lo, fp32 accumulate): fp32-faithful like the per-layer kernels
//   warps 4-7    mid epilogue, thread = row: tcgen05.ld its 64 accumulator columns, BN + ReLU, hi | lo split, and store
//                the row into the XT2 tiles - the layout the next MMA reads (a row is 128 contiguous bytes per k-block;
//                STS.128 with the swizzle XOR: conflict-free)
//   warps 12-15  final epilogue, thread = channel: max / min over the tile's rows per pool group, then
//                relu(scale * ((scale >= 0 ? max : min) + bias) + shift) -> out
//   warp 2       TMEM allocator (512 columns: 2 x 64 acc2 | 2 x 128 acc3 | W2 hi 64 | W2 lo 64)
//
// MODE 1 (train mode: p2c_sa_xyz_linear for a 64 -> 64 second layer whose rows are kept, sa1.1 of the backbone).  The batch
// statistics of conv1's output are what the NEXT kernel needs, so the stack stops after conv1 - and everything per row
// happens in TENSOR MEMORY (640 threads, two transform groups):
//   transform   thread = row: [dx dy dz 1] as tf32 hi | lo -> tcgen05.st (A operand of conv0); later conv0's accumulator
//   (2 x 4 w.)  row comes back (tcgen05.ld), ReLU, hi | lo, tcgen05.st again (A operand of conv1).  No shared memory.
//   MMA warp    MMA1(t): D1 = [d | 1] [A | c]^T (conv0 with its BatchNorm folded into the coefficient tile, K = 8);
//               L2(t-1): acc2 = X1 W1^T + 1 b1^T (K = 64 + one k-step of a column of ones against the bias tile); both TS
//               form with ROWS as the UMMA M and the 64 channels as N: half the tensor time of the channels-as-lanes kernel
//               (linear_tc.cu pads 64 channels to 128 lanes), B operands (8 KB tiles) resident in shared memory
//   epilogue 1  thread = row: accumulator row -> [128 rows x 64 channels] staging tile in the SWIZZLE_128B layout
//   epilogue 2  one lane issues two TMA stores (the tensor map un-swizzles); 128 threads read the tile COLUMN-wise
//               (thread = channel x row half, conflict-free) for the BatchNorm sum / sum of squares - the cross-row
//               reduction that the channels-as-lanes form gets for free
// How it got here (config 2, 1,048,576 rows; channels-as-lanes kernel: 101 us): rows-as-lanes with both operands from shared
// memory 115-135 us; A operand in tensor memory but conv0 evaluated by the transform threads from coefficients in shared
// memory 92 us (ncu: the transform warps wait for their buffer, the epilogue's bias LDS stalls - 64 broadcast LDS.128 per
// row and tile saturate the shared-memory data path, 4 cycles each); conv0 and the bias moved onto the tensor core: 69 us
// (60 us without the store and the statistics; HBM floor 42 us).
#include <cstdlib>

#include "bn_fold.cuh"
#include "tc_common.cuh"

using namespace p2c_tc;

namespace {

constexpr int ST_THREADS = 512;
constexpr int ST_C0 = 64, ST_C1 = 64;               // widths of conv0 / conv1 = K of the two tensor-core layers
constexpr int ST_KB = 2;                            // k-blocks of 32 per layer
constexpr int ST_XT1 = 3, ST_XT2 = 2;               // ring stages (one k-block, hi | lo = 32 KB each)
constexpr uint32_t ST_STAGE = 2u * RAW_BYTES;
constexpr uint32_t ST_W1_TILE = ST_C1 * TC_BK * 4;  // 8 KB: [64 channels][32 floats]
// kind::tf32, fp32 accumulate, K-major A and B, M = 128 rows, N = 64 channels
constexpr uint32_t ST_IDESC_L2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ST_C1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

constexpr uint32_t ST_W0T_OFF = 0;                  // MODE 1, inside the XT1 region: conv0 coefficient tile (hi | lo)
constexpr uint32_t ST_B1T_OFF = 2 * ST_W1_TILE;     // MODE 1: conv1 bias tile (hi | lo)
// MODE 1 tensor-memory columns: acc2 2 x 64 | XA 2 x 144 (hi 64 | ones 8 | lo 64 | pad 8) | DV 2 x 16 (hi 8 | lo 8) | D1 64
constexpr uint32_t ST_XA_COLS = 144, ST_XA_ONE = 64, ST_XA_LO = 72;
constexpr uint32_t ST_TM_XA = 128, ST_TM_DV = ST_TM_XA + 2 * ST_XA_COLS, ST_TM_D1 = ST_TM_DV + 32;
static_assert(ST_TM_D1 + 64 == 512, "MODE 1 tensor-memory budget");
constexpr uint32_t ST_XT1_OFF = 0;
constexpr uint32_t ST_XT2_OFF = ST_XT1_OFF + ST_XT1 * ST_STAGE;
constexpr uint32_t ST_W1_OFF = ST_XT2_OFF + ST_XT2 * ST_STAGE;        // [kb][hi | lo]
constexpr uint32_t ST_D_OFF = ST_W1_OFF + ST_KB * 2 * ST_W1_TILE;     // [buf][row] centred coordinates
constexpr uint32_t ST_W0_OFF = ST_D_OFF + 2 * TC_BM * 16;             // packed conv0 coefficients
constexpr uint32_t ST_SC2_OFF = ST_W0_OFF + ST_C0 * 16;
constexpr uint32_t ST_SH2_OFF = ST_SC2_OFF + ST_C1 * 4;
constexpr uint32_t ST_SC3_OFF = ST_SH2_OFF + ST_C1 * 4;
constexpr uint32_t ST_SH3_OFF = ST_SC3_OFF + 128 * 4;
constexpr uint32_t ST_BAR_OFF = ST_SH3_OFF + 128 * 4;
constexpr uint32_t ST_SMEM = ST_BAR_OFF + 256 + 1024;

struct StArgs {
  P2cXyzFirst g;
  const float* W1; const float* b1;
  const float* W2; const float* b2;
  BnFoldDev bn0, bn1, bn2;
  float* out; int64_t ldo;
  int M, C2, pool, m_tiles;
  // MODE 1
  const float* scale0; const float* shift0;   // conv0's folded BatchNorm as arrays (instead of bn0)
  double* stats;                                // 2 x 64 float64 sums of conv1's raw output, accumulated
};

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

// MODE 1 runs a second group of four transform warps (warps 16-19: 640 threads, 94 registers): group g owns the tiles
// t = g (mod 2) and the A-operand buffer g in tensor memory
template <int MODE>
__global__ void __launch_bounds__(MODE == 1 ? ST_THREADS + 128 : ST_THREADS, 1)
sa_stack_kernel(const __grid_constant__ CUtensorMap tmY, const StArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* xt1_sm = smem + ST_XT1_OFF;
  uint8_t* xt2_sm = smem + ST_XT2_OFF;
  uint8_t* w1_sm = smem + ST_W1_OFF;
  float4* s_d = reinterpret_cast<float4*>(smem + ST_D_OFF);
  float4* s_w0 = reinterpret_cast<float4*>(smem + ST_W0_OFF);
  float* s_sc2 = reinterpret_cast<float*>(smem + ST_SC2_OFF);
  float* s_sh2 = reinterpret_cast<float*>(smem + ST_SH2_OFF);
  float* s_sc3 = reinterpret_cast<float*>(smem + ST_SC3_OFF);
  float* s_sh3 = reinterpret_cast<float*>(smem + ST_SH3_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ST_BAR_OFF);
  uint64_t* xt1_full = bars;                    // [3]
  uint64_t* xt1_empty = xt1_full + ST_XT1;      // [3]
  uint64_t* xt2_full = xt1_empty + ST_XT1;      // [2]
  uint64_t* xt2_empty = xt2_full + ST_XT2;      // [2]
  uint64_t* acc2_full = xt2_empty + ST_XT2;     // [2]
  uint64_t* acc2_empty = acc2_full + 2;
  uint64_t* acc3_full = acc2_empty + 2;
  uint64_t* acc3_empty = acc3_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc3_empty + 2);
  uint64_t* d1_full = bars + 24;                // [2] MODE 1: conv0's accumulator holds a tile of transform group g (a group sees
                                                // only every other tile, so each group needs its OWN phase sequence)

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const bool writer = blockIdx.x == 0;

  // ---- one-time setup ----
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < ST_XT1; ++s) { mbar_init(&xt1_full[s], 4); mbar_init(&xt1_empty[s], 1); }
    for (int s = 0; s < ST_XT2; ++s) { mbar_init(&xt2_full[s], 4); mbar_init(&xt2_empty[s], MODE == 1 ? 4 : 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc2_full[s], 1); mbar_init(&acc2_empty[s], 4);
      mbar_init(&acc3_full[s], 1); mbar_init(&acc3_empty[s], 4);
    }
    mbar_init(&d1_full[0], 1); mbar_init(&d1_full[1], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  if (tid < ST_C0) {
    // conv0 with its BatchNorm folded in: x = max(A . d + c, 0), A = scale w, c = scale b + shift; stored as the packed
    // operands of the transform (see linear_tc.cu): per four channels [Ax0 Ax1 Ay0 Ay1 | Az0 Az1 c0 c1 | Ax2 ... ]
    const int k = tid;
    float sc, sh;
    if (MODE == 1 && a.bn0.active && a.g.moments && a.bn0.stats) {
      // train mode: the first conv's batch statistics in closed form from the nine coordinate moments (bn_fold.cuh)
      double sum, sumsq;
      p2c_xyz_first_sums(a.g.moments, a.bn0.count, (double)__ldg(a.g.W0 + (size_t)k * a.g.ldw0),
                         (double)__ldg(a.g.W0 + (size_t)k * a.g.ldw0 + 1), (double)__ldg(a.g.W0 + (size_t)k * a.g.ldw0 + 2),
                         a.g.b0 ? (double)__ldg(a.g.b0 + k) : 0.0, sum, sumsq);
      if (writer) {
        const_cast<double*>(a.bn0.stats)[k] = sum;
        const_cast<double*>(a.bn0.stats)[a.bn0.C + k] = sumsq;
      }
      p2c_bn_fold_sums(a.bn0, k, writer, sum, sumsq, sc, sh);
    } else if (MODE == 1 && !a.bn0.active) {
      sc = __ldg(a.scale0 + k); sh = __ldg(a.shift0 + k);
    } else
    p2c_bn_fold_channel(a.bn0, k, writer, sc, sh);
    const float* w = a.g.W0 + (size_t)k * a.g.ldw0;
    const float b = a.g.b0 ? __ldg(a.g.b0 + k) : 0.f;
    if (MODE == 1) {
      // conv0 runs on the tensor core too (below): its coefficients [Ax Ay Az c 0 0 0 0] are row k of a K-major
      // SWIZZLE_128B B-operand tile (hi | lo) in the otherwise unused XT1 region; the tile behind it holds conv1's bias as
      // the B operand of one more k-step against a column of ones
      const float v[4] = {sc * __ldg(w), sc * __ldg(w + 1), sc * __ldg(w + 2), fmaf(sc, b, sh)};
      float4 h4, l4;
      float* hp = &h4.x; float* lp = &l4.x;
#pragma unroll
      for (int i = 0; i < 4; ++i) { hp[i] = __uint_as_float(__float_as_uint(v[i]) & 0xffffe000u); lp[i] = v[i] - hp[i]; }
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      uint8_t* t0 = xt1_sm + ST_W0T_OFF + k * 128;
      const uint32_t c0 = ((0u ^ (uint32_t)(k & 7)) << 4), c1 = ((1u ^ (uint32_t)(k & 7)) << 4);
      *reinterpret_cast<float4*>(t0 + c0) = h4;                  *reinterpret_cast<float4*>(t0 + c1) = z4;
      *reinterpret_cast<float4*>(t0 + ST_W1_TILE + c0) = l4;     *reinterpret_cast<float4*>(t0 + ST_W1_TILE + c1) = z4;
      const float b1v = a.b1 ? __ldg(a.b1 + k) : 0.f;
      const float b1h = __uint_as_float(__float_as_uint(b1v) & 0xffffe000u);
      uint8_t* tb = xt1_sm + ST_B1T_OFF + k * 128;
      *reinterpret_cast<float4*>(tb + c0) = make_float4(b1h, 0.f, 0.f, 0.f);                *reinterpret_cast<float4*>(tb + c1) = z4;
      *reinterpret_cast<float4*>(tb + ST_W1_TILE + c0) = make_float4(b1v - b1h, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(tb + ST_W1_TILE + c1) = z4;
    } else {
    float* q = reinterpret_cast<float*>(s_w0) + (k >> 2) * 16 + ((k >> 1) & 1) * 8 + (k & 1);
    q[0] = sc * __ldg(w); q[2] = sc * __ldg(w + 1); q[4] = sc * __ldg(w + 2); q[6] = fmaf(sc, b, sh);
    }
  } else if (tid < ST_C0 + ST_C1) {
    const int k = tid - ST_C0;
    if (MODE == 1) {
      s_sh2[k] = a.b1 ? __ldg(a.b1 + k) : 0.f;       // the raw output keeps its bias; BatchNorm comes in the consumer
    } else {
      float sc, sh;
      p2c_bn_fold_channel(a.bn1, k, writer, sc, sh);
      s_sc2[k] = sc;
      s_sh2[k] = fmaf(sc, a.b1 ? __ldg(a.b1 + k) : 0.f, sh);
    }
  } else if (MODE == 0 && tid < ST_C0 + ST_C1 + 128) {
    const int n = tid - ST_C0 - ST_C1;
    float sc = 0.f, sh = 0.f;
    if (n < a.C2) p2c_bn_fold_channel(a.bn2, n, writer, sc, sh);
    s_sc3[n] = sc;
    s_sh3[n] = sh;
  }
  // W1 (C1, C0) -> shared memory, tf32 hi | lo, K-major SWIZZLE_128B tiles [kb][hi | lo][64 channels][32 floats]
  for (int e = tid; e < ST_C1 * ST_C0; e += ST_THREADS) {
    const int n = e >> 6, k = e & 63;
    const float w = __ldg(a.W1 + (size_t)n * ST_C0 + k);
    const float h = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
    const uint32_t off = (uint32_t)(n * 128 + ((((k & 31) >> 2) ^ (n & 7)) << 4) + (k & 3) * 4);
    uint8_t* t = w1_sm + (size_t)((k >> 5) * 2) * ST_W1_TILE;
    *reinterpret_cast<float*>(t + off) = h;
    *reinterpret_cast<float*>(t + ST_W1_TILE + off) = w - h;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_acc2 = tmem_base;                 // + ab * 64
  const uint32_t tm_acc3 = tmem_base + 128;           // + ab * 128
  const uint32_t tm_w2 = tmem_base + 384;             // hi [0, 64) | lo [64, 128)
  const uint32_t tm_xa = tmem_base + ST_TM_XA;        // MODE 1: conv0's output as the A operand of L2, [buf] x ST_XA_COLS
  const uint32_t tm_dv = tmem_base + ST_TM_DV;        // MODE 1: [d | 1] as the A operand of conv0, [buf] x (hi 8 | lo 8)
  const uint32_t tm_d1 = tmem_base + ST_TM_D1;        // MODE 1: conv0's accumulator (single)
  if (MODE == 1 && ((warp >= 8 && warp < 12))) {
    // the constant column of ones (and its seven zero neighbours) of both XA buffers: A operand of the bias k-step
    const uint32_t one[8] = {__float_as_uint(1.0f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    const uint32_t la = (uint32_t)((warp & 3) * 32) << 16;
    tmem_st8(tm_xa + la + ST_XA_ONE, one);
    tmem_st8(tm_xa + ST_XA_COLS + la + ST_XA_ONE, one);
    tmem_wait_st();
  }

  // W2 (C2, C1) -> tensor memory (thread = output channel; hi | lo), zero padded beyond C2
  if (MODE == 0 && warp >= 8 && warp < 12) {
    const int q = warp & 3;
    const int n = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    for (int kb = 0; kb < ST_KB; ++kb) {
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float w = n < a.C2 ? __ldg(a.W2 + (size_t)n * ST_C1 + kb * TC_BK + c) : 0.f;
        const uint32_t h = __float_as_uint(w) & 0xffffe000u;
        hi[c] = h;
        lo[c] = __float_as_uint(w - __uint_as_float(h));
      }
      tmem_st32(tm_w2 + lane_addr + (uint32_t)(kb * TC_BK), hi);
      tmem_st32(tm_w2 + lane_addr + (uint32_t)(ST_C1 + kb * TC_BK), lo);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int my_tiles = (a.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 1) {
    // ===== MMA issuer (warp-uniform loop, one elected lane issues): L2(0); then L2(t+1), L3(t) =====
    int xs = 0; uint32_t xph = 0;
    auto issue_l2 = [&](int t) {
      const int ab = t & 1;
      mbar_wait(&acc2_empty[ab], (((uint32_t)(t >> 1)) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d = tm_acc2 + (uint32_t)ab * ST_C1;
      for (int kb = 0; kb < ST_KB; ++kb) {
        mbar_wait(&xt1_full[xs], xph);
        tc_fence_after();
        const uint32_t x_hi = smem_u32(xt1_sm + (size_t)xs * ST_STAGE);
        const uint32_t w_hi = smem_u32(w1_sm + (size_t)(kb * 2) * ST_W1_TILE);
        const uint64_t ahi = make_kmajor_sw128_desc(x_hi), alo = make_kmajor_sw128_desc(x_hi + RAW_BYTES);
        const uint64_t bhi = make_kmajor_sw128_desc(w_hi), blo = make_kmajor_sw128_desc(w_hi + ST_W1_TILE);
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma_tf32_ss(d, ahi + (uint64_t)(ks * 2), bhi + (uint64_t)(ks * 2), ST_IDESC_L2, (kb | ks) != 0);
            umma_tf32_ss(d, alo + (uint64_t)(ks * 2), bhi + (uint64_t)(ks * 2), ST_IDESC_L2, 1u);
            umma_tf32_ss(d, ahi + (uint64_t)(ks * 2), blo + (uint64_t)(ks * 2), ST_IDESC_L2, 1u);
          }
          umma_commit(&xt1_empty[xs]);
          if (kb == ST_KB - 1) umma_commit(&acc2_full[ab]);
        }
        __syncwarp();
        if (++xs == ST_XT1) { xs = 0; xph ^= 1; }
      }
    };
    auto issue_l3 = [&](int t) {
      const int ab = t & 1;
      mbar_wait(&acc3_empty[ab], (((uint32_t)(t >> 1)) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d = tm_acc3 + (uint32_t)ab * TC_BM;
      for (int kb = 0; kb < ST_KB; ++kb) {
        mbar_wait(&xt2_full[kb], (uint32_t)t & 1u);
        tc_fence_after();
        const uint32_t x_hi = smem_u32(xt2_sm + (size_t)kb * ST_STAGE);
        const uint64_t bhi = make_kmajor_sw128_desc(x_hi), blo = make_kmajor_sw128_desc(x_hi + RAW_BYTES);
        const uint32_t w_hi = tm_w2 + (uint32_t)(kb * TC_BK), w_lo = w_hi + (uint32_t)ST_C1;
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma_tf32_ts(d, w_hi + ks * 8u, bhi + (uint64_t)(ks * 2), TC_IDESC, (kb | ks) != 0);
            umma_tf32_ts(d, w_lo + ks * 8u, bhi + (uint64_t)(ks * 2), TC_IDESC, 1u);
            umma_tf32_ts(d, w_hi + ks * 8u, blo + (uint64_t)(ks * 2), TC_IDESC, 1u);
          }
          umma_commit(&xt2_empty[kb]);
          if (kb == ST_KB - 1) umma_commit(&acc3_full[ab]);
        }
        __syncwarp();
      }
    };
    if (MODE == 1) {
      // conv0 AND conv1 on the tensor core, A operands in TENSOR MEMORY (lane = row, written by the transform warps with
      // tcgen05.st), B operands (coefficients, W1, bias) resident in shared memory:
      //   MMA1(t):   D1 = [d | 1] [A | c]^T            K = 8, three tf32 passes
      //   L2(t-1):   acc2 = X1 W1^T + 1 b1^T           K = 64 + the bias k-step against the column of ones
      // Measured on the way here: the SS form of L2 and a transform that read its coefficients from shared memory were both
      // bound by the shared-memory data path (operand reads + 64 broadcast LDS.128 per row), not by the tensor pipe.
      const uint32_t w0t = smem_u32(xt1_sm + ST_W0T_OFF), b1t = smem_u32(xt1_sm + ST_B1T_OFF);
      const uint64_t w0hi = make_kmajor_sw128_desc(w0t), w0lo = make_kmajor_sw128_desc(w0t + ST_W1_TILE);
      const uint64_t b1hi = make_kmajor_sw128_desc(b1t), b1lo = make_kmajor_sw128_desc(b1t + ST_W1_TILE);
      for (int t = 0; t <= my_tiles; ++t) {
        if (t < my_tiles) {
          const int db = t & 1;
          mbar_wait(&acc3_empty[db], ((uint32_t)(t >> 1)) & 1u);            // dv_full[db]
          mbar_wait(&xt1_full[2], ((uint32_t)t & 1u) ^ 1u);                 // d1_empty
          tc_fence_after();
          const uint32_t dv = tm_dv + (uint32_t)db * 16u;
          if (elect_one_sync()) {
            umma_tf32_ts(tm_d1, dv, w0hi, ST_IDESC_L2, 0u);
            umma_tf32_ts(tm_d1, dv + 8u, w0hi, ST_IDESC_L2, 1u);
            umma_tf32_ts(tm_d1, dv, w0lo, ST_IDESC_L2, 1u);
            umma_commit(&acc3_full[db]);                                    // dv_empty[db]
            umma_commit(&d1_full[db]);
          }
          __syncwarp();
        }
        if (t >= 1) {
          const int u = t - 1, ab = u & 1;
          mbar_wait(&acc2_empty[ab], (((uint32_t)(u >> 1)) & 1u) ^ 1u);
          mbar_wait(&xt1_full[ab], ((uint32_t)(u >> 1)) & 1u);              // xa_full[ab]
          tc_fence_after();
          const uint32_t d = tm_acc2 + (uint32_t)ab * ST_C1;
          const uint32_t xa = tm_xa + (uint32_t)ab * ST_XA_COLS;
          if (elect_one_sync()) {
#pragma unroll
            for (int kb = 0; kb < ST_KB; ++kb) {
              const uint32_t w_hi = smem_u32(w1_sm + (size_t)(kb * 2) * ST_W1_TILE);
              const uint64_t bhi = make_kmajor_sw128_desc(w_hi), blo = make_kmajor_sw128_desc(w_hi + ST_W1_TILE);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint32_t x_hi = xa + (uint32_t)(kb * TC_BK + ks * 8), x_lo = x_hi + ST_XA_LO;
                umma_tf32_ts(d, x_hi, bhi + (uint64_t)(ks * 2), ST_IDESC_L2, (kb | ks) != 0);
                umma_tf32_ts(d, x_lo, bhi + (uint64_t)(ks * 2), ST_IDESC_L2, 1u);
                umma_tf32_ts(d, x_hi, blo + (uint64_t)(ks * 2), ST_IDESC_L2, 1u);
              }
            }
            umma_tf32_ts(d, xa + ST_XA_ONE, b1hi, ST_IDESC_L2, 1u);         // + bias: 1 * b_hi + 1 * b_lo (exact operands)
            umma_tf32_ts(d, xa + ST_XA_ONE, b1lo, ST_IDESC_L2, 1u);
            umma_commit(&xt1_empty[ab]);                                    // xa_empty[ab]
            umma_commit(&acc2_full[ab]);
          }
          __syncwarp();
        }
      }
    } else {
      if (my_tiles > 0) issue_l2(0);
      for (int t = 0; t < my_tiles; ++t) {
        if (t + 1 < my_tiles) issue_l2(t + 1);
        issue_l3(t);
      }
    }
  } else if ((warp >= 8 && warp < 12) || warp >= 16) {
    // ===== L1: gather + conv0 + BN + ReLU -> XT1 ring (the xyz-first operand transform of linear_tc.cu) =====
    const int tt = (warp & 3) * 32 + lane;          // row of the tile (MODE 0: tid - 256)
    const int cj = tt & 7, r7 = (tt >> 3) & 7, hf = tt >> 6;
    const uint32_t toff = (uint32_t)(hf * 8192 + r7 * 128 + ((cj ^ r7) << 4));
    const int row0 = hf * 64 + r7;
    const P2cXyzFirst& g = a.g;
    auto tile_row = [&](int t) { return (int64_t)((int)blockIdx.x + t * (int)gridDim.x) * TC_BM + tt; };
    auto ldg_f = [](const float* p) { float v; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; };
    auto load_idx = [&](int t) -> int64_t {
      const int64_t r = tile_row(t);
      int64_t v = 0;
      if (t < my_tiles && r < a.M) asm volatile("ld.global.nc.s64 %0, [%1];" : "=l"(v) : "l"(g.idx + r));
      return v;
    };
    struct Raw6 { float px, py, pz, cx, cy, cz; };
    auto load_pc = [&](int t, int64_t p) -> Raw6 {
      const int64_t r = tile_row(t);
      if (t >= my_tiles || r >= a.M) return Raw6{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const unsigned bs = (unsigned)r / (unsigned)g.ns, b = bs / (unsigned)g.S;
      p = (p < 0 || p >= g.N) ? 0 : p;
      const float* pp = g.xyz + ((size_t)b * g.N + (size_t)p) * 3;
      const float* cc = g.new_xyz + (size_t)bs * 3;
      return Raw6{ldg_f(pp), ldg_f(pp + 1), ldg_f(pp + 2), ldg_f(cc), ldg_f(cc + 1), ldg_f(cc + 2)};
    };
    int xs = 0; uint32_t xph = 0;
    const int tg = (MODE == 1 && warp >= 16) ? 1 : 0;      // transform group (MODE 1): tiles tg, tg + 2, ...
    const int tstep = MODE == 1 ? 2 : 1;
    int64_t i1 = load_idx(tg + tstep);
    Raw6 cur = load_pc(tg, load_idx(tg));
    if (MODE == 1) {
      // thread = row.  [A] its centred coordinates, split into tf32 hi | lo with the constant 1 of the bias column, go into
      // tensor memory as the A operand of conv0; [B] conv0's accumulator row comes back, ReLU, hi | lo split, and goes into
      // tensor memory again as the A operand of conv1.  No shared-memory traffic at all in this role.
      const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
      for (int t = tg; t < my_tiles; t += 2) {
        const int64_t i2 = load_idx(t + 4);
        const Raw6 nxt = load_pc(t + 2, i1);
        const int ab = t & 1;                         // = tg: this group's DV / XA buffers
        const uint32_t ph = ((uint32_t)(t >> 1)) & 1u;
        {
          const float d3[3] = {cur.px - cur.cx, cur.py - cur.cy, cur.pz - cur.cz};
          uint32_t hi[8] = {0u, 0u, 0u, __float_as_uint(1.0f), 0u, 0u, 0u, 0u}, lo[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            hi[i] = __float_as_uint(d3[i]) & 0xffffe000u;
            lo[i] = __float_as_uint(d3[i] - __uint_as_float(hi[i]));
          }
          mbar_wait(&acc3_full[ab], ph ^ 1u);                                // dv_empty[ab]
          tc_fence_after();
          tmem_st8(tm_dv + (uint32_t)ab * 16u + lane_addr, hi);
          tmem_st8(tm_dv + (uint32_t)ab * 16u + 8u + lane_addr, lo);
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc3_empty[ab]);                       // dv_full[ab]
        }
        mbar_wait(&d1_full[ab], ph);
        mbar_wait(&xt1_empty[ab], ph ^ 1u);                                  // xa_empty[ab]
        tc_fence_after();
        const uint32_t xa = tm_xa + (uint32_t)ab * ST_XA_COLS + lane_addr;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t raw[16], hi[16], lo[16];
          tmem_ld16(tm_d1 + lane_addr + (uint32_t)(c * 16), raw);
          tmem_wait_ld();
          if (c == 3) {                                                      // D1 is read: conv0 of the next tile may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&xt1_full[2]);                        // d1_empty
          }
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float x0 = fmaxf(__uint_as_float(raw[j]), 0.f), x1 = fmaxf(__uint_as_float(raw[j + 1]), 0.f);
            const uint32_t h0 = __float_as_uint(x0) & 0xffffe000u, h1 = __float_as_uint(x1) & 0xffffe000u;
            const float2 l = __fadd2_rn(make_float2(x0, x1), make_float2(-__uint_as_float(h0), -__uint_as_float(h1)));
            hi[j] = h0; hi[j + 1] = h1;
            lo[j] = __float_as_uint(l.x); lo[j + 1] = __float_as_uint(l.y);
          }
          tmem_st16(xa + (uint32_t)(c * 16), hi);
          tmem_st16(xa + ST_XA_LO + (uint32_t)(c * 16), lo);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&xt1_full[ab]);                           // xa_full[ab]
        cur = nxt;
        i1 = i2;
      }
    } else
    for (int t = 0; t < my_tiles; ++t) {
      const int64_t i2 = load_idx(t + 2);
      const Raw6 nxt = load_pc(t + 1, i1);
      float4* dbuf = s_d + (t & 1) * TC_BM;
      dbuf[tt] = make_float4(cur.px - cur.cx, cur.py - cur.cy, cur.pz - cur.cz, 0.f);
      asm volatile("bar.sync 2, 128;" ::: "memory");
      float4 d[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = dbuf[row0 + 8 * i];
      for (int kb = 0; kb < ST_KB; ++kb) {
        const float4* wq = s_w0 + (kb * (TC_BK / 4) + cj) * 4;
        const float4 p0 = wq[0], p1 = wq[1], p2 = wq[2], p3 = wq[3];
        const float2 ax01 = make_float2(p0.x, p0.y), ay01 = make_float2(p0.z, p0.w), az01 = make_float2(p1.x, p1.y),
                     c01 = make_float2(p1.z, p1.w);
        const float2 ax23 = make_float2(p2.x, p2.y), ay23 = make_float2(p2.z, p2.w), az23 = make_float2(p3.x, p3.y),
                     c23 = make_float2(p3.z, p3.w);
        mbar_wait(&xt1_empty[xs], xph ^ 1);
        uint8_t* hip = xt1_sm + (size_t)xs * ST_STAGE + toff;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 dx = make_float2(d[i].x, d[i].x), dy = make_float2(d[i].y, d[i].y), dz = make_float2(d[i].z, d[i].z);
          const float2 y01 = __ffma2_rn(az01, dz, __ffma2_rn(ay01, dy, __ffma2_rn(ax01, dx, c01)));
          const float2 y23 = __ffma2_rn(az23, dz, __ffma2_rn(ay23, dy, __ffma2_rn(ax23, dx, c23)));
          const float4 x = make_float4(fmaxf(y01.x, 0.f), fmaxf(y01.y, 0.f), fmaxf(y23.x, 0.f), fmaxf(y23.y, 0.f));
          float4 h;
          h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
          h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
          h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
          h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
          const float2 l0 = __fadd2_rn(make_float2(x.x, x.y), make_float2(-h.x, -h.y));
          const float2 l1 = __fadd2_rn(make_float2(x.z, x.w), make_float2(-h.z, -h.w));
          *reinterpret_cast<float4*>(hip + i * 1024) = h;
          *reinterpret_cast<float4*>(hip + RAW_BYTES + i * 1024) = make_float4(l0.x, l0.y, l1.x, l1.y);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&xt1_full[xs]);
        if (++xs == ST_XT1) { xs = 0; xph ^= 1; }
      }
      cur = nxt;
      i1 = i2;
    }
  } else if (MODE == 1 && warp >= 4 && warp < 8) {
    // ===== MODE 1 epilogue, part 1 (thread = row): raw output row + bias -> swizzled staging tile =====
    const int q = warp & 3;
    const int tt = tid - 128;                          // 0..127 = row of the tile
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t roff = (uint32_t)tt * 128u;
    const uint32_t rx = (uint32_t)(tt & 7);
    for (int t = 0; t < my_tiles; ++t) {
      const int ab = t & 1;
      uint8_t* sbuf = xt2_sm + (size_t)(t & 1) * ST_STAGE;       // [kb][128 rows][128 B], SWIZZLE_128B
      // the staging buffers are handed over with NAMED barriers (producer: bar.arrive, consumer: bar.sync, 256 threads; ids
      // 4 / 5 = "buffer b holds tile t", 6 / 7 = "buffer b is free: the statistics warps have read it and its bulk store has
      // drained it") - both sides touch the tile with ordinary loads / stores, and named barriers are what racecheck follows
      if (t >= 2) asm volatile("bar.sync %0, 256;" ::"r"(6 + (t & 1)) : "memory");
      mbar_wait(&acc2_full[ab], ((uint32_t)(t >> 1)) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int kb = 0; kb < ST_KB; ++kb) {
        uint32_t raw[32];
        tmem_ld32(tm_acc2 + (uint32_t)ab * ST_C1 + (uint32_t)(kb * TC_BK) + lane_addr, raw);
        tmem_wait_ld();
        if (kb == ST_KB - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc2_empty[ab]);
        }
        uint8_t* dst = sbuf + (size_t)kb * RAW_BYTES + roff;
#pragma unroll
        for (int c = 0; c < 8; ++c)                  // (the bias is already in the accumulator: the MMA's last k-step)
          *reinterpret_cast<float4*>(dst + (((uint32_t)c ^ rx) << 4)) =
              make_float4(__uint_as_float(raw[4 * c]), __uint_as_float(raw[4 * c + 1]), __uint_as_float(raw[4 * c + 2]),
                          __uint_as_float(raw[4 * c + 3]));
      }
      fence_proxy_async();
      asm volatile("bar.arrive %0, 256;" ::"r"(4 + (t & 1)) : "memory");
    }
  } else if (MODE == 1 && warp >= 12) {
    // ===== MODE 1 epilogue, part 2: TMA store of the staged tile; BatchNorm sums column-wise (thread = channel x half) =====
    const int tt = tid - 384;                          // 0..127
    const int sc_c = tt & 63, sh_h = tt >> 6;          // channel, row half
    const uint32_t soff = (uint32_t)((sc_c >> 5) * RAW_BYTES + (sc_c & 3) * 4);
    const uint32_t scq = (uint32_t)((sc_c & 31) >> 2);
    float h1 = 0.f, l1 = 0.f, h2 = 0.f, l2 = 0.f;      // float-float running sums (see linear_tc.cu)
    auto two_sum = [](float& hi, float& lo, float t) {
      const float s_ = __fadd_rn(hi, t);
      const float bb = __fadd_rn(s_, -hi);
      lo = __fadd_rn(lo, __fadd_rn(__fadd_rn(hi, -__fadd_rn(s_, -bb)), __fadd_rn(t, -bb)));
      hi = s_;
    };
    for (int t = 0; t < my_tiles; ++t) {
      const int m0 = ((int)blockIdx.x + t * (int)gridDim.x) * TC_BM;
      uint8_t* sbuf = xt2_sm + (size_t)(t & 1) * ST_STAGE;
      asm volatile("bar.sync %0, 256;" ::"r"(4 + (t & 1)) : "memory");
      if (tt == 0) {
#pragma unroll
        for (int kb = 0; kb < ST_KB; ++kb)
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmY)), "r"(smem_u32(sbuf + (size_t)kb * RAW_BYTES)), "r"(kb * TC_BK), "r"(m0)
                       : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (a.stats) {
        const int r_lo = sh_h * 64;
        const int valid = min(64, a.M - m0 - r_lo);    // rows past M hold relu(c) of a zero offset: not statistics
        float t1 = 0.f, t2 = 0.f;
        const uint8_t* src = sbuf + soff + (size_t)r_lo * 128;
        if (valid == 64) {
          float2 p1[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, p2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
          for (int i = 0; i < 64; i += 2) {
            const float v0 = *reinterpret_cast<const float*>(src + i * 128 + ((scq ^ (uint32_t)(i & 7)) << 4));
            const float v1 = *reinterpret_cast<const float*>(src + (i + 1) * 128 + ((scq ^ (uint32_t)((i + 1) & 7)) << 4));
            const float2 y = make_float2(v0, v1);
            p1[(i >> 1) & 1] = __fadd2_rn(p1[(i >> 1) & 1], y);
            p2[(i >> 1) & 1] = __ffma2_rn(y, y, p2[(i >> 1) & 1]);
          }
          t1 = (p1[0].x + p1[0].y) + (p1[1].x + p1[1].y);
          t2 = (p2[0].x + p2[0].y) + (p2[1].x + p2[1].y);
        } else {
          for (int i = 0; i < valid; ++i) {
            const float v = *reinterpret_cast<const float*>(src + i * 128 + ((scq ^ (uint32_t)(i & 7)) << 4));
            t1 += v;
            t2 = fmaf(v, v, t2);
          }
        }
        two_sum(h1, l1, t1);
        two_sum(h2, l2, t2);
      }
      // the buffer is free once the bulk store has read it (issued by thread 0 of this role) and every warp here is done
      if (tt == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      if (t + 2 < my_tiles) asm volatile("bar.arrive %0, 256;" ::"r"(6 + (t & 1)) : "memory");   // (no arrival nobody waits for)
    }
    if (tt == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (a.stats) {
      atomicAdd(a.stats + sc_c, (double)h1 + (double)l1);
      atomicAdd(a.stats + ST_C1 + sc_c, (double)h2 + (double)l2);
    }
  } else if (MODE == 0 && warp >= 4 && warp < 8) {
    // ===== mid epilogue: thread = row of the tile; acc2 -> BN + ReLU -> hi | lo -> XT2 (operand of L3) =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t roff = (uint32_t)row * 128u;
    const uint32_t rx = (uint32_t)(row & 7);
    for (int t = 0; t < my_tiles; ++t) {
      const int ab = t & 1;
      mbar_wait(&acc2_full[ab], ((uint32_t)(t >> 1)) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int kb = 0; kb < ST_KB; ++kb) {
        uint32_t raw[32];
        tmem_ld32(tm_acc2 + (uint32_t)ab * ST_C1 + (uint32_t)(kb * TC_BK) + lane_addr, raw);
        tmem_wait_ld();
        if (kb == ST_KB - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc2_empty[ab]);
        }
        mbar_wait(&xt2_empty[kb], ((uint32_t)t & 1u) ^ 1u);
        uint8_t* hip = xt2_sm + (size_t)kb * ST_STAGE + roff;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 sc = *reinterpret_cast<const float4*>(s_sc2 + kb * TC_BK + c * 4);
          const float4 sh = *reinterpret_cast<const float4*>(s_sh2 + kb * TC_BK + c * 4);
          const float2 p0 = __ffma2_rn(make_float2(__uint_as_float(raw[4 * c]), __uint_as_float(raw[4 * c + 1])),
                                       make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
          const float2 p1 = __ffma2_rn(make_float2(__uint_as_float(raw[4 * c + 2]), __uint_as_float(raw[4 * c + 3])),
                                       make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
          const float4 x = make_float4(fmaxf(p0.x, 0.f), fmaxf(p0.y, 0.f), fmaxf(p1.x, 0.f), fmaxf(p1.y, 0.f));
          float4 h;
          h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
          h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
          h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
          h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
          const float2 l0 = __fadd2_rn(make_float2(x.x, x.y), make_float2(-h.x, -h.y));
          const float2 l1 = __fadd2_rn(make_float2(x.z, x.w), make_float2(-h.z, -h.w));
          const uint32_t co = ((uint32_t)c ^ rx) << 4;
          *reinterpret_cast<float4*>(hip + co) = h;
          *reinterpret_cast<float4*>(hip + RAW_BYTES + co) = make_float4(l0.x, l0.y, l1.x, l1.y);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&xt2_full[kb]);
      }
    }
  } else if (MODE == 0 && warp >= 12) {
    // ===== final epilogue: thread = output channel; pool over the rows of each group, then BN + ReLU =====
    const int q = warp & 3;
    const int n = q * 32 + lane;
    const bool n_ok = n < a.C2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float sc = s_sc3[n], sh = s_sh3[n];
    const float bias = (a.b2 && n_ok) ? __ldg(a.b2 + n) : 0.f;
    const float NEG_INF = -__int_as_float(0x7f800000), POS_INF = __int_as_float(0x7f800000);
    const int G = a.pool;
    const int gshift = 31 - __clz(G);
    const int groups = a.M >> gshift;
    float gmx = NEG_INF, gmn = POS_INF;
    for (int t = 0; t < my_tiles; ++t) {
      const int ab = t & 1;
      const int m0 = ((int)blockIdx.x + t * (int)gridDim.x) * TC_BM;
      mbar_wait(&acc3_full[ab], ((uint32_t)(t >> 1)) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t raw[32];
        tmem_ld32(tm_acc3 + (uint32_t)ab * TC_BM + (uint32_t)c * 32u + lane_addr, raw);
        tmem_wait_ld();
        if (c == 3) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc3_empty[ab]);
        }
        float mx[4] = {NEG_INF, NEG_INF, NEG_INF, NEG_INF}, mn[4] = {POS_INF, POS_INF, POS_INF, POS_INF};
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          mx[j & 3] = fmaxf(mx[j & 3], __uint_as_float(raw[j]));
          mn[j & 3] = fminf(mn[j & 3], __uint_as_float(raw[j]));
        }
        gmx = fmaxf(gmx, fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])));
        gmn = fminf(gmn, fminf(fminf(mn[0], mn[1]), fminf(mn[2], mn[3])));
        const int rows_done = c * 32 + 32;
        if ((rows_done & (G - 1)) == 0) {                // G is 32, 64 or 128; tiles start on group boundaries
          const int grp = (m0 + rows_done - G) >> gshift;
          if (n_ok && grp < groups) {
            // BatchNorm is affine and monotone per channel: max over the group commutes with it up to the sign of scale
            const float v = (sc >= 0.f ? gmx : gmn) + bias;
            a.out[(size_t)grp * a.ldo + n] = fmaxf(fmaf(v, sc, sh), 0.f);
          }
          gmx = NEG_INF; gmn = POS_INF;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// p2c_sa_xyz_linear for a 64 -> 64 second layer with its raw output kept (sa1.1 in train mode): MODE 1 of the kernel
// above.  P2C_EUNSUPPORTED = shape not taken (the caller runs the channels-as-lanes kernel of linear_tc.cu).
int p2c_sa_pair_tc(const P2cXyzFirst& g, int B, const float* scale0, const float* shift0, const p2c_bn_fold* bn0,
                   const float* W1, const float* b1, int C0, int N1, float* Y, int64_t ldy, double* stats,
                   cudaStream_t st) {
  if (C0 != ST_C0 || N1 != ST_C1 || !Y) return P2C_EUNSUPPORTED;
  if (!bn0 && !scale0) return P2C_EUNSUPPORTED;                    // a first conv without BatchNorm + ReLU: other kernel
  if ((ldy % 4) != 0 || (reinterpret_cast<uintptr_t>(Y) & 15) != 0) return P2C_EUNSUPPORTED;
  const int64_t rows = (int64_t)B * g.S * g.ns;
  if (rows > 0x7fffffff) return P2C_EUNSUPPORTED;
  CUtensorMap tmY;
  if (int rc = make_map_2d(&tmY, Y, (uint64_t)N1, (uint64_t)rows, (uint64_t)ldy, TC_BK, TC_BM, CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  StArgs a{};
  a.g = g;
  a.W1 = W1; a.b1 = b1;
  a.bn0 = p2c_bn_fold_dev(bn0);
  a.scale0 = scale0; a.shift0 = shift0;
  a.M = (int)rows; a.m_tiles = (int)((rows + TC_BM - 1) / TC_BM);
  a.stats = stats;
  a.pool = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  static int sms_of[64] = {0};
  if (dev < 64 && sms_of[dev] == 0) {
    P2C_CUDA_TRY(cudaFuncSetAttribute(sa_stack_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms_of[dev] = n;
  }
  const int sms = p2c_sm_budget(dev < 64 ? sms_of[dev] : 148);
  const int grid = a.m_tiles < sms ? a.m_tiles : sms;
  sa_stack_kernel<1><<<grid, ST_THREADS + 128, ST_SMEM, st>>>(tmY, a);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

// See include/point2cyl.h
extern "C" int p2c_sa_stack_fused(const float* xyz, const float* new_xyz, const int64_t* idx, int B, int N, int S,
                                  int nsample, const float* W0, int64_t ldw0, const float* b0, const p2c_bn_fold* bn0,
                                  const float* W1, const float* b1, const p2c_bn_fold* bn1, const float* W2,
                                  const float* b2, const p2c_bn_fold* bn2, int C0, int C1, int C2, float* out,
                                  int64_t ldo, void* stream) {
  if (!xyz || !new_xyz || !idx || !W0 || !W1 || !W2 || !bn0 || !bn1 || !bn2 || !out) return P2C_EINVAL;
  if (B <= 0 || N <= 0 || S <= 0 || nsample <= 0 || ldw0 < 3 || ldo < C2) return P2C_EINVAL;
  if (C0 != ST_C0 || C1 != ST_C1 || C2 <= 0 || C2 > 128) return P2C_EUNSUPPORTED;
  if (nsample != 32 && nsample != 64 && nsample != 128) return P2C_EUNSUPPORTED;
  if ((int64_t)B * S * nsample > 0x7fffffff) return P2C_EUNSUPPORTED;
  if (int e = p2c_bn_fold_check(bn0, C0)) return e;
  if (int e = p2c_bn_fold_check(bn1, C1)) return e;
  if (int e = p2c_bn_fold_check(bn2, C2)) return e;
  const int M = B * S * nsample;
  StArgs a{};
  a.g = P2cXyzFirst{xyz, new_xyz, idx, W0, ldw0, b0, N, S, nsample, nullptr};
  a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2;
  a.bn0 = p2c_bn_fold_dev(bn0); a.bn1 = p2c_bn_fold_dev(bn1); a.bn2 = p2c_bn_fold_dev(bn2);
  a.out = out; a.ldo = ldo; a.M = M; a.C2 = C2; a.pool = nsample; a.m_tiles = (M + TC_BM - 1) / TC_BM;
  int dev = 0;
  cudaGetDevice(&dev);
  static int sms_of[64] = {0};
  if (dev < 64 && sms_of[dev] == 0) {
    P2C_CUDA_TRY(cudaFuncSetAttribute(sa_stack_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms_of[dev] = n;
  }
  const int sms = p2c_sm_budget(dev < 64 ? sms_of[dev] : 148);
  const int grid = a.m_tiles < sms ? a.m_tiles : sms;
  CUtensorMap none{};                    // MODE 0 stores nothing through TMA
  sa_stack_kernel<0><<<grid, ST_THREADS, ST_SMEM, (cudaStream_t)stream>>>(none, a);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
