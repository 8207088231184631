// Debug probe for the tcgen05 Lee kernel: one CTA, one K step (8 cells), dumps the TMA-staged smem
// tile and the TMEM accumulator so descriptor / layout assumptions can be checked on hardware.
#include "../../spatialcore_b200/csrc/lee_tc.cu"
#include "../../spatialcore_b200/csrc/common.cu"
#include <vector>
#include <cstdio>
#include <cmath>
#include <cstdlib>
using namespace sc;

__global__ void __launch_bounds__(192)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
             float* smem_dump /*[ (4+8) KB /4 ]*/, float* tmem_dump /*[128][256]*/, int mode) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  __shared__ uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw = tc_smem_u32(tc_smem);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  unsigned char* ring_ptr = tc_smem + (ring - raw);
  if (threadIdx.x == 0) { tc_mbar_init(&full_bar, 1); tc_mbar_init(&done_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_base_slot)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;
  if (threadIdx.x == 0) printf("tmem base 0x%08x ring 0x%x\n", tmem_d, ring);
  if (warp == 0 && lane == 0) {
    tc_mbar_expect_tx(&full_bar, kTcABytes + kTcBBytes);
    tma_load_3d(ring_ptr, &map_a, 0, 0, 0, &full_bar);
    tma_load_3d(ring_ptr + kTcABytes, &map_b, 0, 0, 0, &full_bar);
  }
  tc_mbar_wait(&full_bar, 0);
  // dump smem (generic proxy read after the barrier)
  for (int i = threadIdx.x; i < (int)((kTcABytes + kTcBBytes) / 4); i += blockDim.x)
    smem_dump[i] = reinterpret_cast<float*>(ring_ptr)[i];
  __syncthreads();
  if (warp == 1 && lane == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint64_t ad = umma_desc_mn_sw128(ring, 1024, 512);
    const uint64_t bd = umma_desc_mn_sw128(ring + kTcABytes, 1024, 512);
    umma_tf32(tmem_d, ad, bd, 0u);
    umma_commit(&done_bar);
  }
  tc_mbar_wait(&done_bar, 0);
  __nanosleep(200000);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (mode == 1 && warp >= 2) {  // overwrite column 0..3 of every lane with a pattern: validates tcgen05.ld addressing
    const int lane_grp = warp & 3;
    const uint32_t taddr = tmem_d + ((uint32_t)(lane_grp * 32) << 16);
    uint32_t a = __float_as_uint(1000.f + lane_grp * 32 + lane), b = __float_as_uint(2000.f), c_ = __float_as_uint(3000.f), d = __float_as_uint(4000.f);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" :: "r"(taddr), "r"(a), "r"(b), "r"(c_), "r"(d) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    const int lane_grp = warp & 3;
    for (int c = 0; c < 8; ++c) {
      uint32_t v[32];
      const uint32_t taddr = tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(c * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
          "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
            "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
            "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
            "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) tmem_dump[(lane_grp * 32 + lane) * 256 + c * 32 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(256) : "memory");
}

int main() {
  const int n = 8, ldt = 256;
  std::vector<float> hA(n * ldt), hB(n * ldt);
  for (int i = 0; i < n; ++i) for (int x = 0; x < ldt; ++x) { hA[i * ldt + x] = (x < 128) ? (float)(i == (x % 8)) : 0.f; hB[i * ldt + x] = (float)(x + 1000 * i); }
  float *dA, *dB, *dS, *dT;
  cudaMalloc(&dA, n * ldt * 4); cudaMalloc(&dB, n * ldt * 4); cudaMalloc(&dS, 12288); cudaMalloc(&dT, 128 * 256 * 4);
  cudaMemcpy(dA, hA.data(), n * ldt * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), n * ldt * 4, cudaMemcpyHostToDevice);
  cudaMemset(dT, 0xFF, 128 * 256 * 4);
  CUtensorMap ma, mb;
  if (make_map(&ma, dA, n, ldt, 4) || make_map(&mb, dB, n, ldt, 8)) { printf("map fail %s\n", sc_last_error()); return 1; }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  int mode = getenv("PROBE_MODE") ? atoi(getenv("PROBE_MODE")) : 0;
  probe_kernel<<<1, 192, 32768>>>(ma, mb, dS, dT, mode);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<float> hS(3072), hT(128 * 256);
  cudaMemcpy(hS.data(), dS, 12288, cudaMemcpyDeviceToHost); cudaMemcpy(hT.data(), dT, 128 * 256 * 4, cudaMemcpyDeviceToHost);
  // smem A: expect atom a (32 genes), row k (cell), chunk c^k swizzle.  print raw first rows
  printf("smem A first 2 rows of atom0 (32 floats each):\n");
  for (int r = 0; r < 2; ++r) { for (int j = 0; j < 32; ++j) printf("%g ", hS[r * 32 + j]); printf("\n"); }
  printf("smem B atom0 row0 / row1:\n");
  for (int r = 0; r < 2; ++r) { for (int j = 0; j < 32; ++j) printf("%g ", hS[1024 + r * 32 + j]); printf("\n"); }
  // expected D[x][y] = sum_i A[i][x] B[i][y] = B[x%8][y] for x<128
  double maxerr = 0; int nz = 0;
  for (int x = 0; x < 128; ++x) for (int y = 0; y < 256; ++y) { double want = (double)(y + 1000 * (x % 8)); double got = hT[x * 256 + y]; if (got != 0) nz++; maxerr = fmax(maxerr, fabs(got - want)); }
  printf("tmem nonzeros %d maxerr %g ; D[0][0..7]:", nz, maxerr);
  for (int y = 0; y < 8; ++y) printf(" %g", hT[y]);
  printf("\nD[1][0..7]:"); for (int y = 0; y < 8; ++y) printf(" %g", hT[256 + y]);
  printf("\nD[9][0..7]:"); for (int y = 0; y < 8; ++y) printf(" %g", hT[9 * 256 + y]);
  printf("\n");
  return 0;
}
