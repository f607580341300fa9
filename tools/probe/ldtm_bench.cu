// probe: tcgen05.ld (TMEM -> registers) latency / throughput, 1..4 warps, x16 / x32 / x64 shapes, no MMA running
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int X>
__device__ __forceinline__ uint32_t ld_sum(uint32_t taddr) {
  uint32_t acc = 0;
  if (X == 32) {
    uint32_t v[32];
    asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),"=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31])
      : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += v[i];
  } else {
    uint32_t v[16];
    asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15])
      : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += v[i];
  }
  return acc;
}
template <int X>
__global__ void k(int nwarps_active, int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps_active) {
    acc += ld_sum<X>(base);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) acc += ld_sum<X>(base + (uint32_t)((i * X) & 127));
    t1 = clock64();
  }
  if (lane == 0 && warp < nwarps_active) out[blockIdx.x * 4 + warp] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(256));
}
int main() {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 8 * 4 * 148); cudaMalloc(&sink, 4);
  const int iters = 1000;
  for (int x : {16, 32})
    for (int nw : {1, 2, 4}) {
      cudaMemset(out, 0, 8 * 4 * 148);
      if (x == 32) k<32><<<148, 128>>>(nw, iters, out, sink); else k<16><<<148, 128>>>(nw, iters, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[4]; cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
      printf("x%d, %d warps: %s  cycles/load (warp0) = %.1f  -> %.1f B/cycle/warp, %.1f B/cycle/SM\n", x, nw,
             cudaGetErrorString(e), (double)h[0] / iters, x * 128.0 / ((double)h[0] / iters),
             nw * x * 128.0 / ((double)h[0] / iters));
    }
  return 0;
}
