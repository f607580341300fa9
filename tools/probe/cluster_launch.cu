// probe: which part of a cluster launch does the runtime reject?
#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k(int* out, int CS) {
  extern __shared__ float sm[];
  __shared__ unsigned slot[8];
  cg::cluster_group c = cg::this_cluster();
  unsigned* d = c.map_shared_rank(&slot[c.block_rank()], (c.block_rank() + 1) % CS);
  c.sync();
  if (threadIdx.x == 0) *d = c.block_rank() + 1;
  c.sync();
  if (threadIdx.x == 0) out[blockIdx.x] = slot[(c.block_rank() + CS - 1) % CS] + (int)sm[0] * 0;
}
template <int MAXT>
void run(int CS, int T, size_t smem, const char* tag) {
  int* out; cudaMalloc(&out, 4 * 64);
  auto kk = k<MAXT>;
  cudaError_t e0 = cudaSuccess;
  if (smem > 48 * 1024) e0 = cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(3 * CS); cfg.blockDim = dim3(T); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e1 = cudaLaunchKernelEx(&cfg, kk, out, CS);
  cudaError_t e2 = cudaDeviceSynchronize();
  int h[64] = {0}; cudaMemcpy(h, out, 4 * 3 * CS, cudaMemcpyDeviceToHost);
  printf("%s CS=%d T=%d smem=%zu: attr=%s launch=%s sync=%s out0=%d out1=%d\n", tag, CS, T, smem, cudaGetErrorString(e0),
         cudaGetErrorString(e1), cudaGetErrorString(e2), h[0], h[1]);
  cudaGetLastError();
}
int main() {
  run<1024>(2, 256, 49152, "a");
  run<1024>(2, 256, 0, "b");
  run<512>(8, 256, 49152, "c");
  run<512>(8, 256, 98304, "d");
  run<256>(4, 256, 1024, "e");
  return 0;
}
