// Throughput probe: FADD vs FADD2 (add.rn.f32x2) vs LDS.128 on sm_100a.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fadd2_probe fadd2_probe.cu && ./fadd2_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_fadd(float* out, int iters, float s) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] += s;
  }
  float r = 0;
  for (int i = 0; i < 16; ++i) r += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void k_fadd2(float* out, int iters, float s) {
  unsigned long long a[8];
  for (int i = 0; i < 8; ++i) a[i] = ((unsigned long long)__float_as_uint((float)threadIdx.x) << 32) | __float_as_uint((float)i);
  unsigned long long b = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(b));
  }
  float r = 0;
  for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)a[i]) + __uint_as_float((unsigned)(a[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// LDS.128 from 4 different rows per warp (8 lanes per 128-byte row piece), as in lag_tile_kernel
__global__ void k_lds(float* out, int iters) {
  __shared__ float4 sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  float4 acc = make_float4(0, 0, 0, 0);
  const int q = threadIdx.x & 7;
  unsigned row = (threadIdx.x >> 3) * 37u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 v = sm[((row + i * 61u) & 255u) * 8 + q];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    row += 97u;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = 148, iters = 4096;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int rep = 0; rep < 2; ++rep) {
    float ms;
    dim3 grid(sms * 4), block(256);
    cudaEventRecord(e0); k_fadd<<<grid, block>>>(out, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double adds = (double)grid.x * 256 * iters * 16;
    printf("FADD : %.3f ms  %.1f Gadd/s  (%.1f adds/clk/SM at %d MHz nominal)\n", ms, adds / ms / 1e6, adds / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
    cudaEventRecord(e0); k_fadd2<<<grid, block>>>(out, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("FADD2: %.3f ms  %.1f Gadd/s  (%.1f adds/clk/SM)\n", ms, adds / ms / 1e6, adds / (ms * 1e-3) / sms / (clk * 1e3));
    cudaEventRecord(e0); k_lds<<<grid, block>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double bytes = (double)grid.x * 256 * iters * 8 * 16;
    printf("LDS.128 + 4 FADD: %.3f ms  %.1f GB/s  (%.1f B/clk/SM)\n", ms, bytes / ms / 1e6, bytes / (ms * 1e-3) / sms / (clk * 1e3));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
