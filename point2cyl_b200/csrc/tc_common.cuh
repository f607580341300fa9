// Shared pieces of the tcgen05 MLP-layer kernels (linear_tc.cu: W resident in TMEM; linear_tc_ss.cu: W streamed
// through shared memory for large K): PTX wrappers, descriptors, the per-channel epilogue chunk.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace p2c_tc {

constexpr int TC_BM = 128;           // rows of X per tile  (UMMA N)
constexpr int TC_BN = 128;           // output channels per CTA (UMMA M, TMEM lanes)
constexpr int TC_BK = 32;            // fp32 elements per k-block = 128 bytes = one SWIZZLE_128B row
constexpr int TC_THREADS = 384;
constexpr int RAW_BYTES = TC_BM * TC_BK * 4;   // 16 KB
constexpr int MAX_RAW = 8, MAX_XT = 4;

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
// One lane of a fully converged warp (elect.sync).  The MMA-issuing role runs WARP-UNIFORM code and elects the issuing
// lane per instruction group: inside an `if (lane == 0)` region the compiler treats every operand of tcgen05.mma /
// tcgen05.commit as potentially divergent and wraps each instruction in a waterfall loop (R2UR + ELECT + BRA.U.ANY,
// ~100 issue cycles per MMA - measured: the issuing warp, not the tensor pipe, paced every layer).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
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
// kind::tf32, fp32 accumulate, A and B K-major, M = 128 channels, N = 128 rows
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BM >> 3) << 17) |
                              ((uint32_t)(TC_BN >> 4) << 24);

// timeline probe: role r (0 mma, 1 transform, 2 epilogue, 3 producer) of CTA 0 appends (tag, globaltimer)
__device__ __forceinline__ void dbg_mark(long long* dbg, int role, int& n, int tag) {
  if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && n < 255) {
    long long c;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(c));
    dbg[role * 512 + 2 * n] = tag;
    dbg[role * 512 + 2 * n + 1] = c;
    ++n;
  }
}

// One 32-row chunk of the epilogue for one output channel.  raw[j] = accumulator of row j.  Everything is
// fully unrolled over registers; FULL = all 32 rows valid (the common case, no predicates).
template <bool FULL>
__device__ __forceinline__ void epi_chunk(const uint32_t (&raw)[32], float bias, int jmax, float* stage,
                                          float* yp, int64_t ldy, bool do_stats, bool do_pool, float& t1,
                                          float& t2, float& mx_out, float& mn_out) {
  // packed fp32 pairs (FADD2 / FFMA2 on sm_100): half the issue slots of the bias add and of the statistics
  float2 v2[16];
  const float2 b2 = make_float2(bias, bias);
#pragma unroll
  for (int i = 0; i < 16; ++i)
    v2[i] = __fadd2_rn(make_float2(__uint_as_float(raw[2 * i]), __uint_as_float(raw[2 * i + 1])), b2);
  if (stage) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {                               // rows past M are clipped by the TMA store
      stage[(2 * i) * 32] = v2[i].x;
      stage[(2 * i + 1) * 32] = v2[i].y;
    }
  } else if (yp) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (FULL || 2 * i < jmax) yp[(size_t)(2 * i) * ldy] = v2[i].x;
      if (FULL || 2 * i + 1 < jmax) yp[(size_t)(2 * i + 1) * ldy] = v2[i].y;
    }
  }
  if (do_stats) {
    float2 p1[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};   // four scalar chains in two packed ones
    float2 p2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float2 y = v2[i];
      if (!FULL) { y.x = (2 * i < jmax) ? y.x : 0.f; y.y = (2 * i + 1 < jmax) ? y.y : 0.f; }
      p1[i & 1] = __fadd2_rn(p1[i & 1], y);
      p2[i & 1] = __ffma2_rn(y, y, p2[i & 1]);
    }
    t1 += (p1[0].x + p1[0].y) + (p1[1].x + p1[1].y);
    t2 += (p2[0].x + p2[0].y) + (p2[1].x + p2[1].y);
  }
  if (do_pool) {
    const float NEG_INF = -__int_as_float(0x7f800000), POS_INF = __int_as_float(0x7f800000);
    float mx[4] = {NEG_INF, NEG_INF, NEG_INF, NEG_INF}, mn[4] = {POS_INF, POS_INF, POS_INF, POS_INF};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (FULL || 2 * i < jmax) { mx[i & 3] = fmaxf(mx[i & 3], v2[i].x); mn[i & 3] = fminf(mn[i & 3], v2[i].x); }
      if (FULL || 2 * i + 1 < jmax) { mx[i & 3] = fmaxf(mx[i & 3], v2[i].y); mn[i & 3] = fminf(mn[i & 3], v2[i].y); }
    }
    mx_out = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
    mn_out = fminf(fminf(mn[0], mn[1]), fminf(mn[2], mn[3]));
  }
}


typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
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

// 2-D fp32 tensor map: dims (cols, rows), row stride ld elements, box (bc, br)
inline int make_map_2d(CUtensorMap* tm, const float* base, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t bc,
                       uint32_t br, CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return (int)cudaErrorNotSupported;
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {ld * 4};
  const cuuint32_t box[2] = {bc, br};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

}  // namespace p2c_tc
