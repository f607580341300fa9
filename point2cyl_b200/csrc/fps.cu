// Farthest point sampling, one CTA per cloud (replaces models/pointnet_util.py:63-84).
//
// The npoint rounds are strictly sequential (each arg-max feeds the next centroid), so the kernel
// is latency bound, not bandwidth bound: the cloud is read from HBM exactly once (12*N bytes),
// lives in shared memory as SoA, and each thread keeps its PPT points plus their running distance
// in registers.  One round = PPT distance updates per thread, a two-instruction warp arg-max
// (redux.sync max on the float bits, redux.sync min on the candidate indices), one __syncthreads
// and a second redux pair over the per-warp winners (double-buffered slots, so one barrier per
// round is enough).  Tie-break and rounding follow the reference bit for bit:
//   dist = (dx*dx + dy*dy) + dz*dz, every op rounded (intrinsics: never contracted to FMA);
//   running = dist < running ? dist : running (init 1e10); farthest = FIRST index of the max.
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

template <int PPT, bool XYZ_IN_REGS, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
fps_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ start, int N, int npoint,
           int64_t* __restrict__ out_idx, float* __restrict__ out_xyz) {
  extern __shared__ float s_xyz[];  // xs[N] ys[N] zs[N]
  __shared__ unsigned s_val[2][32];
  __shared__ unsigned s_idx[2][32];
  const int T = blockDim.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int nwarps = T >> 5;
  const int b = blockIdx.x;
  float* xs = s_xyz;
  float* ys = s_xyz + N;
  float* zs = s_xyz + 2 * N;

  const float* p = xyz + (size_t)b * N * 3;
  for (int i = tid; i < 3 * N; i += T) {
    float v = __ldg(p + i);
    int pt = i / 3;
    int c = i - pt * 3;
    s_xyz[c * N + pt] = v;
  }
  __syncthreads();

  float px[XYZ_IN_REGS ? PPT : 1], py[XYZ_IN_REGS ? PPT : 1], pz[XYZ_IN_REGS ? PPT : 1];
  float run[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    int i = j * T + tid;
    bool ok = i < N;
    // padded slots hold running distance 0: they can only tie, and ties go to the lowest index
    run[j] = ok ? 1e10f : 0.0f;
    if (XYZ_IN_REGS) {
      px[j] = ok ? xs[i] : 0.0f;
      py[j] = ok ? ys[i] : 0.0f;
      pz[j] = ok ? zs[i] : 0.0f;
    }
  }

  int far = (int)start[b];
  int buf = 0;
  int64_t* oidx = out_idx + (size_t)b * npoint;
  float* oxyz = out_xyz + (size_t)b * npoint * 3;

  for (int it = 0; it < npoint; ++it) {
    const float cx = xs[far], cy = ys[far], cz = zs[far];
    if (tid == 0) {
      oidx[it] = far;
      oxyz[it * 3 + 0] = cx;
      oxyz[it * 3 + 1] = cy;
      oxyz[it * 3 + 2] = cz;
    }
    float best = 0.0f;
    if (XYZ_IN_REGS && (PPT % 2 == 0)) {
      // two points per instruction: FADD2 / FMUL2 round each half exactly like the scalar ops (round-to-nearest,
      // no contraction), so the distances are bit-identical while the issue slots per point drop from 10 to 6
      const float2 ncx = make_float2(-cx, -cx), ncy = make_float2(-cy, -cy), ncz = make_float2(-cz, -cz);
#pragma unroll
      for (int j = 0; j < PPT; j += 2) {
        const float2 dx = __fadd2_rn(make_float2(px[j], px[j + 1]), ncx);
        const float2 dy = __fadd2_rn(make_float2(py[j], py[j + 1]), ncy);
        const float2 dz = __fadd2_rn(make_float2(pz[j], pz[j + 1]), ncz);
        const float2 d = __fadd2_rn(__fadd2_rn(__fmul2_rn(dx, dx), __fmul2_rn(dy, dy)), __fmul2_rn(dz, dz));
        run[j] = fminf(run[j], d.x);
        run[j + 1] = fminf(run[j + 1], d.y);
        best = fmaxf(best, fmaxf(run[j], run[j + 1]));
      }
    } else {
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      float x, y, z;
      if (XYZ_IN_REGS) {
        x = px[j]; y = py[j]; z = pz[j];
      } else {
        int i = j * T + tid;
        i = i < N ? i : N - 1;
        x = xs[i]; y = ys[i]; z = zs[i];
      }
      float dx = __fsub_rn(x, cx), dy = __fsub_rn(y, cy), dz = __fsub_rn(z, cz);
      float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (!XYZ_IN_REGS) d = (j * T + tid < N) ? d : 0.0f;
      run[j] = fminf(run[j], d);
      best = fmaxf(best, run[j]);
    }
    }
    // distances are >= +0, so their bit patterns order like unsigned integers
    const unsigned vb = __float_as_uint(best);
    const unsigned wmax = __reduce_max_sync(P2C_FULL_MASK, vb);
    unsigned mine = 0xffffffffu;
    if (vb == wmax) {
#pragma unroll
      for (int j = PPT - 1; j >= 0; --j)
        if (__float_as_uint(run[j]) == wmax) mine = (unsigned)(j * T + tid);
    }
    const unsigned widx = __reduce_min_sync(P2C_FULL_MASK, mine);
    if (lane == 0) {
      s_val[buf][warp] = wmax;
      s_idx[buf][warp] = widx;
    }
    __syncthreads();
    const unsigned v = lane < nwarps ? s_val[buf][lane] : 0u;
    const unsigned ix = lane < nwarps ? s_idx[buf][lane] : 0xffffffffu;
    const unsigned gmax = __reduce_max_sync(P2C_FULL_MASK, v);
    far = (int)__reduce_min_sync(P2C_FULL_MASK, v == gmax ? ix : 0xffffffffu);
    buf ^= 1;
  }
}

// Cluster variant for clouds that do not fit one CTA (N > 16384) - and optionally for smaller ones: CS CTAs of a
// thread-block cluster share one cloud, each owns a contiguous slice of ceil(N/CS) points (coordinates and running
// distances in registers, the slice also in shared memory as SoA).  Per round every CTA reduces its slice to one
// candidate (distance bits, global index, x, y, z), thread 0 stores it into slot [round & 1][rank] of EVERY CTA of
// the cluster through distributed shared memory, one cluster barrier publishes the CS candidates, and all CTAs pick
// the same winner (max distance, lowest index) - whose coordinates travel with the candidate, so no CTA ever needs
// a point outside its slice.  Same arithmetic and tie-break as the single-CTA kernel: bit-exact.
constexpr int FPS_MAX_CS = 8;

template <int PPT, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
fps_cluster_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ start, int N, int npoint, int CS,
                   int slice, int64_t* __restrict__ out_idx, float* __restrict__ out_xyz) {
  extern __shared__ float s_xyz[];  // xs[slice] ys[slice] zs[slice]
  __shared__ unsigned s_val[2][32];
  __shared__ unsigned s_idx[2][32];
  __shared__ unsigned s_cand[2][FPS_MAX_CS][5];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int T = blockDim.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int nwarps = T >> 5;
  const int b = blockIdx.x / CS;
  const int base = rank * slice;                       // first global point index of this CTA's slice
  const int cnt = max(0, min(slice, N - base));        // points in the slice
  float* xs = s_xyz;
  float* ys = s_xyz + slice;
  float* zs = s_xyz + 2 * slice;
  const float* p = xyz + ((size_t)b * N + base) * 3;
  for (int i = tid; i < 3 * cnt; i += T) {
    const float v = __ldg(p + i);
    const int pt = i / 3;
    s_xyz[(i - pt * 3) * slice + pt] = v;
  }
  __syncthreads();
  float px[PPT], py[PPT], pz[PPT], run[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int i = j * T + tid;
    const bool ok = i < cnt;
    run[j] = ok ? 1e10f : 0.0f;
    px[j] = ok ? xs[i] : 0.0f;
    py[j] = ok ? ys[i] : 0.0f;
    pz[j] = ok ? zs[i] : 0.0f;
  }
  int far = (int)start[b];
  const float* fp = xyz + ((size_t)b * N + far) * 3;
  float cx = __ldg(fp), cy = __ldg(fp + 1), cz = __ldg(fp + 2);
  int64_t* oidx = out_idx + (size_t)b * npoint;
  float* oxyz = out_xyz + (size_t)b * npoint * 3;
  cluster.sync();                                      // every CTA of the cluster is resident before remote stores
  int buf = 0;
  for (int it = 0; it < npoint; ++it) {
    if (rank == 0 && tid == 0) {
      oidx[it] = far;
      oxyz[it * 3 + 0] = cx;
      oxyz[it * 3 + 1] = cy;
      oxyz[it * 3 + 2] = cz;
    }
    float best = 0.0f;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const float dx = __fsub_rn(px[j], cx), dy = __fsub_rn(py[j], cy), dz = __fsub_rn(pz[j], cz);
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      run[j] = fminf(run[j], d);                       // padded slots stay 0
      best = fmaxf(best, run[j]);
    }
    const unsigned vb = __float_as_uint(best);
    const unsigned wmax = __reduce_max_sync(P2C_FULL_MASK, vb);
    unsigned mine = 0xffffffffu;
    if (vb == wmax) {
#pragma unroll
      for (int j = PPT - 1; j >= 0; --j)
        if (__float_as_uint(run[j]) == wmax && j * T + tid < cnt) mine = (unsigned)(j * T + tid);
    }
    const unsigned widx = __reduce_min_sync(P2C_FULL_MASK, mine);
    if (lane == 0) {
      s_val[buf][warp] = wmax;
      s_idx[buf][warp] = widx;
    }
    __syncthreads();
    if (warp == 0) {
      const unsigned v = lane < nwarps ? s_val[buf][lane] : 0u;
      const unsigned ix = lane < nwarps ? s_idx[buf][lane] : 0xffffffffu;
      const unsigned gmax = __reduce_max_sync(P2C_FULL_MASK, v);
      const unsigned gidx = __reduce_min_sync(P2C_FULL_MASK, (v == gmax) ? ix : 0xffffffffu);
      // lanes 0..CS-1 each publish the candidate to one CTA of the cluster (an empty slice publishes "no candidate")
      if (lane < CS) {
        unsigned* dst = cluster.map_shared_rank(&s_cand[buf][rank][0], lane);
        const bool has = gidx != 0xffffffffu;
        dst[0] = has ? gmax : 0u;
        dst[1] = has ? (unsigned)(base + (int)gidx) : 0xffffffffu;
        dst[2] = has ? __float_as_uint(xs[gidx]) : 0u;
        dst[3] = has ? __float_as_uint(ys[gidx]) : 0u;
        dst[4] = has ? __float_as_uint(zs[gidx]) : 0u;
      }
    }
    cluster.sync();
    unsigned bv = 0u, bi = 0xffffffffu;
    int br = 0;
    for (int r = 0; r < CS; ++r) {
      const unsigned v = s_cand[buf][r][0], ix = s_cand[buf][r][1];
      if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; br = r; }
    }
    far = (int)bi;
    cx = __uint_as_float(s_cand[buf][br][2]);
    cy = __uint_as_float(s_cand[buf][br][3]);
    cz = __uint_as_float(s_cand[buf][br][4]);
    buf ^= 1;
  }
  cluster.sync();                                      // no CTA exits while peers may still store into its slots
}

template <int PPT, int MAXT>
int launch_fps_cluster(const float* xyz, const int64_t* start, int B, int N, int npoint, int64_t* out_idx,
                       float* out_xyz, int CS, cudaStream_t st) {
  auto round32 = [](int v) { return (v + 31) / 32 * 32; };
  const int slice = (N + CS - 1) / CS;
  const int T = round32((slice + PPT - 1) / PPT);
  if (T > MAXT || CS > FPS_MAX_CS) return P2C_EUNSUPPORTED;
  const size_t smem = (size_t)slice * 3 * sizeof(float);
  auto k = fps_cluster_kernel<PPT, MAXT>;
  // static + dynamic shared memory above 48 KB needs the opt-in (the kernels hold < 2 KB of static arrays)
  if (smem + 2048 > 48 * 1024) P2C_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * CS));
  cfg.blockDim = dim3((unsigned)T);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  P2C_CUDA_TRY(cudaLaunchKernelEx(&cfg, k, xyz, start, N, npoint, CS, slice, out_idx, out_xyz));
  return 0;
}

template <int PPT, bool R, int MAXT>
int launch_fps(const float* xyz, const int64_t* start, int B, int N, int npoint, int64_t* out_idx,
               float* out_xyz, int T, cudaStream_t st) {
  size_t smem = (size_t)N * 3 * sizeof(float);
  auto k = fps_kernel<PPT, R, MAXT>;
  // static + dynamic shared memory above 48 KB needs the opt-in (the kernels hold < 2 KB of static arrays)
  if (smem + 2048 > 48 * 1024) P2C_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<B, T, smem, st>>>(xyz, start, N, npoint, out_idx, out_xyz);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

}  // namespace

extern "C" int p2c_fps(const float* xyz, const int64_t* start, int B, int N, int npoint,
                       int64_t* out_idx, float* out_xyz, void* stream) {
  if (!xyz || !start || !out_idx || !out_xyz || B <= 0 || N <= 0 || npoint <= 0) return P2C_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  auto round32 = [](int v) { return (v + 31) / 32 * 32; };
  const char* ev = getenv("P2C_FPS_PPT");   // tools only: force points-per-thread (8, 16 or 32)
  const int force = ev ? atoi(ev) : 0;
  const char* ec = getenv("P2C_FPS_CLUSTER");   // tools only: force the cluster kernel with this many CTAs per cloud
  const int fcs = ec ? atoi(ec) : 0;
  if (fcs == 2 || fcs == 4 || fcs == 8)
    return launch_fps_cluster<16, 1024>(xyz, start, B, N, npoint, out_idx, out_xyz, fcs, st);
  if (force == 16 && N <= 8192)
    return launch_fps<16, true, 512>(xyz, start, B, N, npoint, out_idx, out_xyz, round32((N + 15) / 16), st);
  if (force == 32 && N <= 8192)
    return launch_fps<32, true, 256>(xyz, start, B, N, npoint, out_idx, out_xyz, round32((N + 31) / 32), st);
  if (N <= 1024) {
    int T = round32((N + 3) / 4);
    return launch_fps<4, true, 256>(xyz, start, B, N, npoint, out_idx, out_xyz, T, st);
  }
  if (N <= 4096) {
    int T = round32((N + 7) / 8);
    return launch_fps<8, true, 1024>(xyz, start, B, N, npoint, out_idx, out_xyz, T, st);
  }
  if (N <= 8192)   // measured on B200 at N=8192: 32 points/thread (8 warps) 329 us, 16: 332 us, 8: 375 us per 512 rounds
    return launch_fps<32, true, 256>(xyz, start, B, N, npoint, out_idx, out_xyz, round32((N + 31) / 32), st);
  if (N <= 16384) return launch_fps<16, false, 1024>(xyz, start, B, N, npoint, out_idx, out_xyz, 1024, st);
  if (N <= 65536) return launch_fps_cluster<16, 512>(xyz, start, B, N, npoint, out_idx, out_xyz, 8, st);
  if (N <= 131072) return launch_fps_cluster<16, 1024>(xyz, start, B, N, npoint, out_idx, out_xyz, 8, st);
  return P2C_EUNSUPPORTED;
}
