// Weight gradient of a 1x1-conv layer on the tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   dW[n, k] += sum_m dY[m, n] * A[m, k],   A = X | max(X*scale+shift, 0),      db[n] += sum_m dY[m, n]
//
// The reduction runs over the ROW dimension m (up to 10^6 rows) and the result is a small (N x K) matrix, so the
// work is split over the rows: CTA x owns a contiguous slab of rows and one 128 x 128 tile of dW, accumulates it
// in tensor memory and adds it to global memory once at the end (fp32 red.global) - a split-K GEMM whose "K" is m.
//
// Both operands are row matrices with the channel dimension contiguous, i.e. for the UMMA they are MN-MAJOR:
//   A operand  = dY^T   (M_umma = n, 128 channels),   B operand = A (N_umma = k, 128 channels),   K_umma = rows.
// For 32-bit (tf32) MN-major operands the tensor core accepts exactly one shared-memory layout, the "128-byte
// swizzle with 32-byte atoms" (UMMA layout type SWIZZLE_128B_BASE32B: rows of 128 B, the 32-byte chunk index XORed
// with row & 3, atoms of 4 rows = 512 B).  TMA produces it natively (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): a box of
// [32 rows x 32 channels] lands as 8 such atoms 512 B apart along K (= SBO), and the four 32-channel groups of an
// operand are one box = 4096 B apart along MN (= LBO), so no transposition is ever done: the transform warps only apply the folded BatchNorm +
// ReLU of the previous layer to the X boxes and split both operands into tf32 hi / lo halves in place-shaped
// buffers.  3xTF32 like the forward: dYhi*Ahi + dYlo*Ahi + dYhi*Alo, fp32 accumulate (fp32-faithful).
// kind::tf32 reads the upper 19 bits of an fp32 word and ignores the low 13 mantissa bits, so the RAW fp32 tile IS
// the "hi" operand: only lo = x - trunc(x) is written by the transform (and Ahi when the operand load folds a
// BatchNorm+ReLU - written IN PLACE over the RAW X boxes: every transform thread overwrites exactly the words it read,
// so no second buffer is needed).  That keeps four 32 KB TMA stages in flight per SM and three transformed stages
// ahead of the tensor core; a RAW stage is released by the MMA's commit, not by the transform.
//
// Warp roles (384 threads, 1 CTA/SM): warp 0 TMA producer (8 boxes = 32 KB per 32-row stage), warp 1 MMA issuer,
// warp 2 TMEM allocator, warps 4-11 operand transform, warps 4-7 afterwards the epilogue (thread = dW row n).
#include "tc_common.cuh"

using namespace p2c_tc;

namespace {

constexpr int WG_THREADS = 384;
constexpr int WG_ROWS = 32;                  // rows per stage = 4 UMMA k-steps of 8 rows
constexpr int WG_BOX = WG_ROWS * 128;        // bytes of one [32 rows x 32 channels] box
constexpr int WG_OP = 4 * WG_BOX;            // one operand stage: 128 channels = 16 KB
constexpr int WG_RAW_STAGE = 2 * WG_OP;      // dY | X      (TMA destination; also the tf32 "hi" operands, see below)
constexpr int WG_XT_STAGE = 2 * WG_OP;       // dYlo | Alo
constexpr int WG_RAW = 4, WG_XT = 3;         // 7 x 32 KB: four TMA stages in flight, three transformed stages ahead of the MMAs
constexpr int WG_TILE = 128;

// kind::tf32, fp32 accumulate, A and B MN-major, M = 128, N = 128
constexpr uint32_t WG_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                              ((uint32_t)(WG_TILE >> 3) << 17) | ((uint32_t)(WG_TILE >> 4) << 24);

constexpr uint32_t WG_IDESC_N64 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                  ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(WG_TILE >> 4) << 24);

// MN-major SWIZZLE_128B_BASE32B descriptor: LBO = stride between 32-channel groups, SBO = stride between 4-row atoms
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(WG_BOX >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;                  // descriptor version (sm_100)
  d |= (uint64_t)1 << 61;                  // SWIZZLE_128B_BASE32B
  return d;
}
// byte offset of the 16-byte chunk cj (4 channels) of row r inside a [32 rows x 128 B] box swizzled with 32-byte atoms
__device__ __forceinline__ uint32_t box_off32(int r, int cj) {
  return (uint32_t)r * 128u + ((((uint32_t)(cj >> 1) ^ ((uint32_t)r & 3u)) << 5) | (((uint32_t)cj & 1u) << 4));
}

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

struct WgArgs {
  const float* in_scale; const float* in_shift;
  float* dW; int64_t lddw; float* db;
  int M, N, K;
  int stages_total, stages_per_cta;
  // K <= 64 (the 64-channel layers of sa1: a million rows each): half of a 128-wide X operand would be zero-filled.
  //   fold (N <= 64 too): a stage holds 64 rows - boxes 2, 3 of both operands carry channels 0..63 of the rows
  //     m0+32.., so D[0:64, 0:64] and D[64:128, 64:128] are the products of the two 32-row halves (the off-diagonal
  //     blocks are cross terms nobody reads): the same twelve MMAs reduce twice the rows;
  //   narrow (N > 64): the X operand is two boxes and the MMA shape 128 x 64.
  int fold, narrow;
};

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* raw_sm = smem;                                            // [WG_RAW][dY | X]
  uint8_t* xt_sm = smem + WG_RAW * WG_RAW_STAGE;                     // [WG_XT][dYlo | Alo]
  float* s_scale = reinterpret_cast<float*>(xt_sm + WG_XT * WG_XT_STAGE);
  float* s_shift = s_scale + WG_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + WG_TILE);
  uint64_t* raw_full = bars;                  // [WG_RAW]
  uint64_t* raw_empty = bars + WG_RAW;        // [WG_RAW]
  uint64_t* xt_full = bars + 2 * WG_RAW;      // [WG_XT]
  uint64_t* xt_empty = xt_full + WG_XT;       // [WG_XT]
  uint64_t* acc_full = xt_empty + WG_XT;      // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * WG_TILE, k0 = blockIdx.z * WG_TILE;
  const int st0 = blockIdx.x * a.stages_per_cta;
  const int st1 = min(a.stages_total, st0 + a.stages_per_cta);
  const int my_stages = st1 - st0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmDY)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
    for (int s = 0; s < WG_RAW; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 1); }
    for (int s = 0; s < WG_XT; ++s) { mbar_init(&xt_full[s], 8); mbar_init(&xt_empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, WG_TILE);
  for (int k = tid; k < WG_TILE; k += WG_THREADS) {
    const bool ok = a.in_scale != nullptr && k0 + k < a.K;
    s_scale[k] = ok ? __ldg(a.in_scale + k0 + k) : 0.f;
    s_shift[k] = ok ? __ldg(a.in_shift + k0 + k) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm_acc = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: 4 boxes of dY and 4 boxes of X per stage (warp-uniform loop, elected issuer) =====
    {
      int s = 0; uint32_t ph = 0;
      for (int t = 0; t < my_stages; ++t) {
        const int m0 = (st0 + t) * (a.fold ? 2 * WG_ROWS : WG_ROWS);
        mbar_wait(&raw_empty[s], ph ^ 1);
        if (elect_one_sync()) {
          uint8_t* dst = raw_sm + (size_t)s * WG_RAW_STAGE;
          if (a.fold) {
            mbar_arrive_expect_tx(&raw_full[s], WG_RAW_STAGE);
#pragma unroll
            for (int b = 0; b < 4; ++b)
              tma_load_2d(dst + b * WG_BOX, &tmDY, &raw_full[s], n0 + (b & 1) * 32, m0 + (b >> 1) * WG_ROWS);
#pragma unroll
            for (int b = 0; b < 4; ++b)
              tma_load_2d(dst + WG_OP + b * WG_BOX, &tmX, &raw_full[s], k0 + (b & 1) * 32, m0 + (b >> 1) * WG_ROWS);
          } else {
            mbar_arrive_expect_tx(&raw_full[s], a.narrow ? WG_OP + 2 * WG_BOX : WG_RAW_STAGE);
#pragma unroll
            for (int b = 0; b < 4; ++b) tma_load_2d(dst + b * WG_BOX, &tmDY, &raw_full[s], n0 + b * 32, m0);
#pragma unroll
            for (int b = 0; b < 4; ++b)
              if (b < 2 || !a.narrow) tma_load_2d(dst + WG_OP + b * WG_BOX, &tmX, &raw_full[s], k0 + b * 32, m0);
          }
        }
        __syncwarp();
        if (++s == WG_RAW) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-uniform loop, one elected lane issues (tc_common.cuh: elect_one_sync) =====
    {
      int xs = 0; uint32_t xph = 0;
      int rs = 0;
      const uint32_t idesc = a.narrow ? WG_IDESC_N64 : WG_IDESC;
      for (int t = 0; t < my_stages; ++t) {
        mbar_wait(&xt_full[xs], xph);               // implies raw_full[rs]: the transform read that stage
        tc_fence_after();
        const uint32_t rawb = smem_u32(raw_sm + (size_t)rs * WG_RAW_STAGE);
        const uint32_t xtb = smem_u32(xt_sm + (size_t)xs * WG_XT_STAGE);
        const uint32_t ahib = rawb + WG_OP;           // raw X, or relu(bn(X)) written in place by the transform
        if (elect_one_sync()) {
          // the four 8-row k-steps differ in the descriptors' start-address field only (+1024 B = +64 units)
          const uint64_t dyhi = make_mnmajor_sw128_desc(rawb), dylo = make_mnmajor_sw128_desc(xtb);
          const uint64_t ahi = make_mnmajor_sw128_desc(ahib), alo = make_mnmajor_sw128_desc(xtb + WG_OP);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t o = (uint64_t)(ks * 64);
            umma_tf32_ss(tm_acc, dyhi + o, ahi + o, idesc, (t | ks) != 0);
            umma_tf32_ss(tm_acc, dylo + o, ahi + o, idesc, 1u);
            umma_tf32_ss(tm_acc, dyhi + o, alo + o, idesc, 1u);
          }
          umma_commit(&xt_empty[xs]);
          umma_commit(&raw_empty[rs]);                // the tensor core has read the RAW stage: TMA may refill it
          if (t == my_stages - 1) umma_commit(acc_full);
        }
        __syncwarp();
        if (++xs == WG_XT) { xs = 0; xph ^= 1; }
        if (++rs == WG_RAW) rs = 0;
      }
    }
  } else if (warp >= 4) {
    // ===== operand transform: thread = (box, 16-byte channel chunk cj, 8-row group rg) =====
    const int tt = tid - 128;
    const int box = tt >> 5;                       // 0-3: dY, 4-7: X
    const int cj = tt & 7, rg = (tt >> 3) & 3;
    const bool is_x = box >= 4;
    const bool has_affine = is_x && a.in_scale != nullptr;
    const int cbox = a.fold ? (box & 1) : (box & 3);     // 32-channel group this box carries
    const bool idle = a.narrow && box >= 6;              // narrow: X boxes 2, 3 do not exist (the warp only keeps the barriers going)
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_affine) {
      sc = *reinterpret_cast<const float4*>(s_scale + cbox * 32 + cj * 4);
      sh = *reinterpret_cast<const float4*>(s_shift + cbox * 32 + cj * 4);
    }
    const bool do_bias = !is_x && a.db != nullptr && blockIdx.z == 0;
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t raw_op = is_x ? WG_OP : 0;
    const uint32_t xt_op = is_x ? WG_OP : 0;       // lo half of this operand
    const uint32_t box_off = (uint32_t)(box & 3) * WG_BOX;
    int s = 0; uint32_t ph = 0;
    int xs = 0; uint32_t xph = 0;
    for (int t = 0; t < my_stages; ++t) {
      mbar_wait(&raw_full[s], ph);
      uint8_t* rawp = raw_sm + (size_t)s * WG_RAW_STAGE + raw_op + box_off;
      float4 x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rg * 8 + i;
        x[i] = idle ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(rawp + box_off32(r, cj));
      }
      if (++s == WG_RAW) { s = 0; ph ^= 1; }
      if (has_affine) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          x[i].x = fmaxf(fmaf(x[i].x, sc.x, sh.x), 0.f);
          x[i].y = fmaxf(fmaf(x[i].y, sc.y, sh.y), 0.f);
          x[i].z = fmaxf(fmaf(x[i].z, sc.z, sh.z), 0.f);
          x[i].w = fmaxf(fmaf(x[i].w, sc.w, sh.w), 0.f);
        }
      }
      if (do_bias) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { bsum.x += x[i].x; bsum.y += x[i].y; bsum.z += x[i].z; bsum.w += x[i].w; }
      }
      mbar_wait(&xt_empty[xs], xph ^ 1);
      uint8_t* lop = xt_sm + (size_t)xs * WG_XT_STAGE + xt_op + box_off;
      uint8_t* hip = rawp;                           // Ahi (X boxes with an affine fold only): in place over the RAW box
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rg * 8 + i;
        float4 l;
        l.x = x[i].x - __uint_as_float(__float_as_uint(x[i].x) & 0xffffe000u);
        l.y = x[i].y - __uint_as_float(__float_as_uint(x[i].y) & 0xffffe000u);
        l.z = x[i].z - __uint_as_float(__float_as_uint(x[i].z) & 0xffffe000u);
        l.w = x[i].w - __uint_as_float(__float_as_uint(x[i].w) & 0xffffe000u);
        const uint32_t off = box_off32(r, cj);
        if (idle) continue;
        *reinterpret_cast<float4*>(lop + off) = l;
        if (has_affine) *reinterpret_cast<float4*>(hip + off) = x[i];   // the tensor core drops the low 13 bits itself
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&xt_full[xs]);
      if (++xs == WG_XT) { xs = 0; xph ^= 1; }
    }
    if (do_bias) {
      const int n = n0 + cbox * 32 + cj * 4;
      if (n + 0 < a.N) atomicAdd(a.db + n + 0, bsum.x);
      if (n + 1 < a.N) atomicAdd(a.db + n + 1, bsum.y);
      if (n + 2 < a.N) atomicAdd(a.db + n + 2, bsum.z);
      if (n + 3 < a.N) atomicAdd(a.db + n + 3, bsum.w);
    }
    if (warp < 8) {
      // ===== epilogue: thread = row n of the dW tile; 32 accumulator columns (k) at a time =====
      const int q = warp & 3;
      // fold: lanes 64..127 hold the second 32-row half's product of the same dW rows, in columns 64..127
      const int n = n0 + (a.fold ? (q & 1) : q) * 32 + lane;
      const int c_first = a.fold ? (q >> 1) * 2 : 0, c_count = (a.fold || a.narrow) ? 2 : 4;
      const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
      mbar_wait(acc_full, 0);
      tc_fence_after();
      // every CTA of a slab column adds into the same dW tile: 16-byte vector reductions where the row is aligned (a
      // quarter of the L2 atomic operations), and the CTAs start at different 32-column chunks so they do not all
      // queue on the same addresses at once
#pragma unroll 1
      for (int cc = 0; cc < c_count; ++cc) {
        const int ct = c_first + ((cc + (int)blockIdx.x) & (c_count - 1));   // accumulator chunk (32 columns)
        const int c = a.fold ? (ct & 1) : ct;                                // its 32-column group of dW
        uint32_t raw[32];
        tmem_ld32(tm_acc + lane_addr + (uint32_t)ct * 32u, raw);
        tmem_wait_ld();
        if (n < a.N) {
          float* row = a.dW + (size_t)n * a.lddw + k0 + c * 32;
          const bool vec = (reinterpret_cast<uintptr_t>(row) & 15) == 0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (vec && k0 + c * 32 + j + 3 < a.K) {
              atomicAdd(reinterpret_cast<float4*>(row + j),
                        make_float4(__uint_as_float(raw[j]), __uint_as_float(raw[j + 1]), __uint_as_float(raw[j + 2]),
                                    __uint_as_float(raw[j + 3])));
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (k0 + c * 32 + j + i < a.K) atomicAdd(row + j + i, __uint_as_float(raw[j + i]));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tm_acc, WG_TILE);
  }
}

constexpr int WG_SMEM = WG_RAW * WG_RAW_STAGE + WG_XT * WG_XT_STAGE + 2 * WG_TILE * 4 + 256 + 1024;

}  // namespace

// 1 when the tensor-core kernel takes this call (pure function of the arguments)
extern "C" int p2c_wgrad_path(int64_t lddy, int64_t ldx, int64_t M, int N, int K, int has_mask) {
  if (has_mask) return 0;
  if ((lddy % 4) != 0 || (ldx % 4) != 0) return 0;     // TMA global strides are multiples of 16 bytes
  if (M < 1024) return 0;                              // tiny reductions: launch + prologue cost dominates
  (void)N; (void)K;
  return 1;
}

int p2c_wgrad_tc(const float* dY, int64_t lddy, const float* X, int64_t ldx, const float* in_scale,
                 const float* in_shift, int64_t M, int N, int K, float* dW, int64_t lddw, float* db, cudaStream_t st) {
  if (!p2c_wgrad_path(lddy, ldx, M, N, K, 0)) return P2C_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(dY) & 15) || (reinterpret_cast<uintptr_t>(X) & 15)) return P2C_EUNSUPPORTED;
  if (M > 0x7fffffff) return P2C_EUNSUPPORTED;
  CUtensorMap tmDY, tmX;
  int rc = make_map_2d(&tmDY, dY, (uint64_t)N, (uint64_t)M, (uint64_t)lddy, 32, WG_ROWS,
                       CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  rc = make_map_2d(&tmX, X, (uint64_t)K, (uint64_t)M, (uint64_t)ldx, 32, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  int dev = 0;
  cudaGetDevice(&dev);
  static int sms_of[64] = {0};
  if (dev < 64 && sms_of[dev] == 0) {
    P2C_CUDA_TRY(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms_of[dev] = n;
  }
  const int sms = p2c_sm_budget(dev < 64 ? sms_of[dev] : 148);
  const int n_tiles = (N + WG_TILE - 1) / WG_TILE, k_tiles = (K + WG_TILE - 1) / WG_TILE;
  const int fold = (N <= 64 && K <= 64) ? 1 : 0, narrow = (!fold && K <= 64) ? 1 : 0;
  const int stage_rows = fold ? 2 * WG_ROWS : WG_ROWS;
  const int stages_total = (int)((M + stage_rows - 1) / stage_rows);
  int gx = sms / (n_tiles * k_tiles);
  if (gx < 1) gx = 1;
  // at least 8 stages (256 rows) per CTA so the prologue / epilogue amortise
  const int max_gx = (stages_total + 7) / 8;
  if (gx > max_gx) gx = max_gx;
  const int spc = (stages_total + gx - 1) / gx;
  gx = (stages_total + spc - 1) / spc;
  WgArgs a{in_scale, in_shift, dW, lddw, db, (int)M, N, K, stages_total, spc, fold, narrow};
  dim3 grid(gx, n_tiles, k_tiles);
  wgrad_tc_kernel<<<grid, WG_THREADS, WG_SMEM, st>>>(tmDY, tmX, a);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
