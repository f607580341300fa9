// Ball query and grouping gather (replaces models/pointnet_util.py:87-107 and :124-139 / :146-163).
//
// Ball query: one warp per FOUR query centres (a point's coordinates are loaded from shared memory once and tested
// against four centres held in registers: 0.6 of the instructions per pair of the one-centre-per-warp version), one
// CTA = 32 centres of the same cloud.  The cloud is
// streamed through shared memory in tiles of TILE points as SoA (x, y, z, |p|^2), each warp tests 32
// points per step, ballots, and appends the hits in index order (prefix popcount), so the output is
// "first nsample indices, ascending" without the reference's (B,S,N) distance matrix and full sort.
// A warp stops at nsample hits; the CTA stops when all its warps have.  The membership test follows
// the reference's rounding exactly (p2c_sqdist_expanded in common.cuh).
//
// Algorithmic traffic per cloud: 12*N (xyz) + 12*S (centres) + 8*S*nsample (indices).
#include "common.cuh"

namespace {

constexpr int BQ_WARPS = 8;
constexpr int BQ_TILE = 1024;
constexpr int BQ_C = 4;            // query centres per warp: a point's four shared-memory loads serve four tests

__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int N, int S,
                  float r2, int nsample, int64_t* __restrict__ out) {
  __shared__ float sx[BQ_TILE], sy[BQ_TILE], sz[BQ_TILE], sn[BQ_TILE];
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int q0 = (blockIdx.x * BQ_WARPS + warp) * BQ_C;
  const float* p = xyz + (size_t)b * N * 3;

  // every per-centre value below is warp-uniform (derived from ballots), so the branches on it do not diverge
  float ax[BQ_C], ay[BQ_C], az[BQ_C], na[BQ_C];
  int cnt[BQ_C], first[BQ_C];
  int64_t* o[BQ_C];
#pragma unroll
  for (int c = 0; c < BQ_C; ++c) {
    const bool valid = q0 + c < S;
    ax[c] = ay[c] = az[c] = 0.f;
    if (valid) {
      const float* cp = new_xyz + ((size_t)b * S + q0 + c) * 3;
      ax[c] = __ldg(cp); ay[c] = __ldg(cp + 1); az[c] = __ldg(cp + 2);
    }
    na[c] = p2c_norm2_rn(ax[c], ay[c], az[c]);
    o[c] = out + ((size_t)b * S + (valid ? q0 + c : 0)) * nsample;
    cnt[c] = valid ? 0 : nsample;              // a centre past S counts as finished
    first[c] = N;
  }
  const unsigned lt_mask = (1u << lane) - 1u;
  auto open = [&]() { bool a = false;
#pragma unroll
    for (int c = 0; c < BQ_C; ++c) a = a || cnt[c] < nsample;
    return a; };

  for (int base = 0; base < N; base += BQ_TILE) {
    const int tn = min(BQ_TILE, N - base);
    for (int i = tid; i < tn; i += BQ_WARPS * 32) {
      const float* s = p + (size_t)(base + i) * 3;
      float x = __ldg(s), y = __ldg(s + 1), z = __ldg(s + 2);
      sx[i] = x; sy[i] = y; sz[i] = z;
      sn[i] = p2c_norm2_rn(x, y, z);
    }
    __syncthreads();
    if (open()) {
      for (int j0 = 0; j0 < tn; j0 += 32) {
        const int j = j0 + lane;
        const bool inb = j < tn;
        const float px = inb ? sx[j] : 0.f, py = inb ? sy[j] : 0.f, pz = inb ? sz[j] : 0.f, pn = inb ? sn[j] : 0.f;
        // two centres per packed operation (FMUL2 / FFMA2 / FADD2 round each half like the scalar ops of
        // p2c_sqdist_expanded: bit-identical distances): ((-2 dot) + |c|^2) + |p|^2, dot = fma(cz,pz, fma(cy,py, cx px))
        float dd[BQ_C];
        const float2 px2 = make_float2(px, px), py2 = make_float2(py, py), pz2 = make_float2(pz, pz), pn2 = make_float2(pn, pn);
#pragma unroll
        for (int c = 0; c < BQ_C; c += 2) {
          const float2 dot = __ffma2_rn(make_float2(az[c], az[c + 1]), pz2,
                                        __ffma2_rn(make_float2(ay[c], ay[c + 1]), py2,
                                                   __fmul2_rn(make_float2(ax[c], ax[c + 1]), px2)));
          const float2 d2 = __fadd2_rn(__fadd2_rn(__fmul2_rn(make_float2(-2.0f, -2.0f), dot), make_float2(na[c], na[c + 1])), pn2);
          dd[c] = d2.x; dd[c + 1] = d2.y;
        }
#pragma unroll
        for (int c = 0; c < BQ_C; ++c) {
          if (cnt[c] >= nsample) continue;
          const float d = dd[c];
          const bool in = inb && !(d > r2);
          const unsigned m = __ballot_sync(P2C_FULL_MASK, in);
          if (m) {
            const int pos = cnt[c] + __popc(m & lt_mask);
            if (in && pos < nsample) o[c][pos] = base + j;
            if (first[c] == N) first[c] = base + j0 + __ffs(m) - 1;
            cnt[c] += __popc(m);
          }
        }
        if (!open()) break;
      }
    }
    const int active = __syncthreads_or(open());
    if (!active) break;
  }
#pragma unroll
  for (int c = 0; c < BQ_C; ++c) {
    if (q0 + c >= S) continue;
    const int have = min(cnt[c], nsample);
    for (int j = have + lane; j < nsample; j += 32) o[c][j] = first[c];
  }
}

// One warp per output row r = (b, s, j).  Lanes stride over the 3 + D columns.
__global__ void __launch_bounds__(256)
group_kernel(const float* __restrict__ xyz, const float* __restrict__ feats, int64_t ldf,
             const float* __restrict__ new_xyz, const int64_t* __restrict__ idx, int N, int S,
             int nsample, int D, float* __restrict__ out, int64_t ldo, int64_t rows) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int64_t bs = r / nsample;           // b*S + s
  const int64_t b = bs / S;
  const int64_t src = idx ? idx[r] : (r - bs * nsample);
  float* o = out + r * ldo;
  const int C = 3 + D;
  const bool ok = src >= 0 && src < N;     // a query with no hit yields N: emit zeros, not garbage
  if (lane < 3) {
    float v = 0.f;
    if (ok) {
      v = __ldg(xyz + ((size_t)b * N + src) * 3 + lane);
      if (new_xyz) v -= __ldg(new_xyz + (size_t)bs * 3 + lane);
    }
    o[lane] = v;
  }
  if (D > 0) {
    const float* f = feats + ((size_t)b * N + (ok ? src : 0)) * ldf;
    for (int c = lane; c < D; c += 32) o[3 + c] = ok ? __ldg(f + c) : 0.f;
  }
  for (int c = C + lane; c < ldo; c += 32) o[c] = 0.f;
}

}  // namespace

extern "C" int p2c_ball_query(const float* xyz, const float* new_xyz, int B, int N, int S, float r2,
                              int nsample, int64_t* out_idx, void* stream) {
  if (!xyz || !new_xyz || !out_idx || B <= 0 || N <= 0 || S <= 0 || nsample <= 0) return P2C_EINVAL;
  dim3 grid(p2c_ceil_div(S, BQ_WARPS * BQ_C), B);
  ball_query_kernel<<<grid, BQ_WARPS * 32, 0, (cudaStream_t)stream>>>(xyz, new_xyz, N, S, r2, nsample,
                                                                      out_idx);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_group(const float* xyz, const float* feats, int64_t ldf, const float* new_xyz,
                         const int64_t* idx, int B, int N, int S, int nsample, int D, float* out,
                         int64_t ldo, void* stream) {
  if (!xyz || !out || B <= 0 || N <= 0 || S <= 0 || nsample <= 0 || D < 0) return P2C_EINVAL;
  if (D > 0 && !feats) return P2C_EINVAL;
  if (ldo < 3 + D) return P2C_EINVAL;
  if (!idx && (S != 1 || nsample != N)) return P2C_EINVAL;  // group-all form
  const int64_t rows = (int64_t)B * S * nsample;
  const int wpb = 8;
  group_kernel<<<p2c_ceil_div(rows, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      xyz, feats, ldf, new_xyz, idx, N, S, nsample, D, out, ldo, rows);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
