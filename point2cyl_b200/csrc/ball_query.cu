// Ball query and grouping gather (replaces models/pointnet_util.py:87-107 and :124-139 / :146-163).
//
// Ball query: one warp per query centre, one CTA = WARPS centres of the same cloud.  The cloud is
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

__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int N, int S,
                  float r2, int nsample, int64_t* __restrict__ out) {
  __shared__ float sx[BQ_TILE], sy[BQ_TILE], sz[BQ_TILE], sn[BQ_TILE];
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int q = blockIdx.x * BQ_WARPS + warp;
  const bool valid = q < S;
  const float* p = xyz + (size_t)b * N * 3;

  float ax = 0.f, ay = 0.f, az = 0.f;
  if (valid) {
    const float* c = new_xyz + ((size_t)b * S + q) * 3;
    ax = __ldg(c); ay = __ldg(c + 1); az = __ldg(c + 2);
  }
  const float na = p2c_norm2_rn(ax, ay, az);
  int64_t* o = out + ((size_t)b * S + (valid ? q : 0)) * nsample;
  int cnt = 0;
  int first = N;
  const unsigned lt_mask = (1u << lane) - 1u;

  for (int base = 0; base < N; base += BQ_TILE) {
    const int tn = min(BQ_TILE, N - base);
    for (int i = tid; i < tn; i += BQ_WARPS * 32) {
      const float* s = p + (size_t)(base + i) * 3;
      float x = __ldg(s), y = __ldg(s + 1), z = __ldg(s + 2);
      sx[i] = x; sy[i] = y; sz[i] = z;
      sn[i] = p2c_norm2_rn(x, y, z);
    }
    __syncthreads();
    if (valid && cnt < nsample) {
      for (int j0 = 0; j0 < tn; j0 += 32) {
        const int j = j0 + lane;
        bool in = false;
        if (j < tn) {
          float d = p2c_sqdist_expanded(ax, ay, az, na, sx[j], sy[j], sz[j], sn[j]);
          in = !(d > r2);
        }
        const unsigned m = __ballot_sync(P2C_FULL_MASK, in);
        if (m) {
          const int pos = cnt + __popc(m & lt_mask);
          if (in && pos < nsample) o[pos] = base + j;
          if (first == N) first = base + j0 + __ffs(m) - 1;
          cnt += __popc(m);
          if (cnt >= nsample) break;
        }
      }
    }
    const int active = __syncthreads_or(valid && cnt < nsample);
    if (!active) break;
  }
  if (valid) {
    const int have = min(cnt, nsample);
    for (int j = have + lane; j < nsample; j += 32) o[j] = first;
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
  dim3 grid(p2c_ceil_div(S, BQ_WARPS), B);
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
