// Does a (1,2,1) cluster launch work on this box, with and without large dynamic shared memory?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(320, 1) k(int* out) {
  extern __shared__ unsigned char sm[];
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) { sm[0] = 1; atomicAdd(out, (int)r + sm[0] - 1); }
}
static void go(dim3 grid, dim3 cl, size_t dyn) {
  int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  cudaLaunchConfig_t cfg = {}; cfg.gridDim = grid; cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = dyn;
  cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension;
  a[0].val.clusterDim.x = cl.x; a[0].val.clusterDim.y = cl.y; a[0].val.clusterDim.z = cl.z;
  cfg.attrs = a; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, d);
  cudaError_t e2 = cudaDeviceSynchronize();
  int h = -1; cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
  printf("grid %u,%u,%u cluster %u,%u,%u dyn %zu: launch %s, sync %s, sum of ranks %d\n", grid.x, grid.y, grid.z, cl.x, cl.y, cl.z, dyn,
         cudaGetErrorString(e), cudaGetErrorString(e2), h);
  cudaGetLastError(); cudaFree(d);
}
int main() {
  go(dim3(1, 2, 64), dim3(1, 2, 1), 1024);
  go(dim3(1, 2, 64), dim3(1, 2, 1), 164864);
  go(dim3(2, 1, 64), dim3(2, 1, 1), 164864);
  go(dim3(4, 8, 9), dim3(1, 2, 1), 164864);
  go(dim3(4, 8, 9), dim3(1, 2, 1), 100000);
  return 0;
}
