// Function-level point-set helpers kept for API completeness of the drop-in module
// (square_distance, models/pointnet_util.py:19-40; index_points, :43-60).  The fused pipeline never
// materialises a (B,S,N) distance matrix; these exist because callers of the reference may.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
square_distance_kernel(const float* __restrict__ src, const float* __restrict__ dst, int S, int N,
                       float* __restrict__ out) {
  const int b = blockIdx.z;
  const int s = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const float* a = src + ((size_t)b * S + s) * 3;
  const float ax = __ldg(a), ay = __ldg(a + 1), az = __ldg(a + 2);
  const float na = p2c_norm2_rn(ax, ay, az);
  if (n >= N) return;
  const float* p = dst + ((size_t)b * N + n) * 3;
  const float bx = __ldg(p), by = __ldg(p + 1), bz = __ldg(p + 2);
  out[((size_t)b * S + s) * N + n] = p2c_sqdist_expanded(ax, ay, az, na, bx, by, bz, p2c_norm2_rn(bx, by, bz));
}

// out[b, m, :] = points[b, idx[b, m], :]; one warp per output row
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ points, int64_t ldp, const int64_t* __restrict__ idx, int N,
                   int Mper, int C, float* __restrict__ out, int64_t rows) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int64_t b = r / Mper;
  int64_t src = idx[r];
  src = src < 0 ? src + N : src;
  const bool ok = src >= 0 && src < N;
  const float* p = points + ((size_t)b * N + (ok ? src : 0)) * ldp;
  float* o = out + r * C;
  for (int c = lane; c < C; c += 32) o[c] = ok ? __ldg(p + c) : 0.f;
}

}  // namespace

extern "C" int p2c_square_distance(const float* src, const float* dst, int B, int S, int N, float* out,
                                   void* stream) {
  if (!src || !dst || !out || B <= 0 || S <= 0 || N <= 0 || S > 65535 || B > 65535) return P2C_EINVAL;
  dim3 grid(p2c_ceil_div(N, 256), S, B);
  square_distance_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, S, N, out);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}

extern "C" int p2c_gather_rows(const float* points, int64_t ldp, const int64_t* idx, int B, int N,
                               int Mper, int C, float* out, void* stream) {
  if (!points || !idx || !out || B <= 0 || N <= 0 || Mper <= 0 || C <= 0 || ldp < C) return P2C_EINVAL;
  const int64_t rows = (int64_t)B * Mper;
  gather_rows_kernel<<<p2c_ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(points, ldp, idx, N, Mper, C, out, rows);
  P2C_RETURN_IF_CUDA_ERROR();
  return 0;
}
