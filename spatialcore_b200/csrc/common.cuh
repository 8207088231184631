// Shared helpers for libsc_b200 (sm_100a).  Internal header, not part of the C ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sc_b200.h"

namespace sc {

void set_error(const char* fmt, ...);

#define SC_CHECK_ARG(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      sc::set_error(__VA_ARGS__);    \
      return SC_ERR_INVALID;         \
    }                                \
  } while (0)

#define SC_CUDA_OK(expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      sc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                    __LINE__);                                                        \
      return SC_ERR_CUDA;                                                             \
    }                                                                                 \
  } while (0)

void count_launch();

// after every kernel launch: count it (sc_launch_count) and surface launch errors
#define SC_LAUNCH_OK()     \
  do {                     \
    sc::count_launch();    \
    SC_CUDA_OK(cudaGetLastError()); \
  } while (0)

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace (256-byte aligned slices).
struct Arena {
  char* base;
  size_t cap;
  size_t off;
  Arena(void* p, size_t bytes) : base(static_cast<char*>(p)), cap(bytes), off(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    if (off + bytes > cap) return nullptr;
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
};

// Number of SMs of the current device (cached per thread).
int sm_count();

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
// Streaming 128-bit load: read once, do not keep in L1.
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ double shfl_f64(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(kFull, lo, src);
  hi = __shfl_sync(kFull, hi, src);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_up_f64(double v, int d) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_up_sync(kFull, lo, d);
  hi = __shfl_up_sync(kFull, hi, d);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_xor_f64(double v, int mask) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(kFull, lo, mask);
  hi = __shfl_xor_sync(kFull, hi, mask);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(kFull, lo, o);
    hi = __shfl_xor_sync(kFull, hi, o);
    v += __hiloint2double(hi, lo);
  }
  return v;
}

// ---- shared-memory barriers (mbarrier) -------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
#endif

// ---- Philox4x32-10 and the keyed Feistel bijection used for on-the-fly permutations ----------
// Mirrored bit-for-bit by spatialcore_b200/philox.py (host) so tests can replay device
// permutations on the CPU oracle.

constexpr int kFeistelRounds = 8;

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                       uint32_t c3, uint32_t k0, uint32_t k1,
                                                       uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Round keys of permutation `perm_index` under `seed`.
__host__ __device__ __forceinline__ void perm_round_keys(uint64_t seed, uint64_t perm_index,
                                                         uint32_t keys[kFeistelRounds]) {
#pragma unroll
  for (int b = 0; b < kFeistelRounds / 4; ++b) {
    uint32_t o[4];
    philox4x32_10((uint32_t)perm_index, (uint32_t)(perm_index >> 32), (uint32_t)b, 0x5C0B200u,
                  (uint32_t)seed, (uint32_t)(seed >> 32), o);
    keys[4 * b + 0] = o[0]; keys[4 * b + 1] = o[1];
    keys[4 * b + 2] = o[2]; keys[4 * b + 3] = o[3];
  }
}

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu;
  x ^= x >> 13; x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}

// Generalised (group-operation) Feistel network on Z_a x Z_b with b = 2^s ~ sqrt(n) and
// a = ceil(n / b): the domain a*b exceeds n by less than b, so cycle-walking into [0, n) almost
// never iterates (acceptance >= 1 - b/n) and warps do not diverge.  State (L, R); even rounds map
// (Z_a, Z_b) -> (Z_b, Z_a) with R' = (L + F(R)) mod a, odd rounds map back with R' = (L + F(R)) mod b.
// kFeistelRounds (even) rounds; a bijection of [0, a*b) for every key set.
struct PermDomain {
  uint32_t n;
  uint32_t a;  // size of the "high" factor
  uint32_t s;  // log2 of the "low" factor b
};

__host__ __device__ __forceinline__ PermDomain make_perm_domain(uint32_t n) {
  uint32_t bits = 1;
  while (bits < 32 && (1ull << bits) < (uint64_t)n) ++bits;  // bits = ceil(log2 n), >= 1
  PermDomain d;
  d.n = n;
  d.s = (bits + 1) / 2;
  d.a = (uint32_t)(((uint64_t)n + (1ull << d.s) - 1) >> d.s);
  if (d.a == 0) d.a = 1;
  return d;
}

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t x, uint32_t y) {
  return (uint32_t)(((uint64_t)x * (uint64_t)y) >> 32);
}

__host__ __device__ __forceinline__ uint32_t feistel_once(uint32_t v, const PermDomain& d,
                                                          const uint32_t* keys) {
  const uint32_t bmask = (1u << d.s) - 1u;
  uint32_t L = v >> d.s, R = v & bmask;  // L in Z_a, R in Z_b
#pragma unroll
  for (int r = 0; r < kFeistelRounds; r += 2) {
    // (Z_a, Z_b) -> (Z_b, Z_a)
    uint32_t f = mulhi32(mix32(R * 0x9E3779B1u + keys[r]), d.a);  // uniform in [0, a)
    uint32_t t = L + f;
    if (t >= d.a) t -= d.a;
    L = R;
    R = t;
    // (Z_b, Z_a) -> (Z_a, Z_b)
    uint32_t f2 = mix32(R * 0x9E3779B1u + keys[r + 1]) >> (32u - d.s);  // uniform in [0, b)
    uint32_t t2 = (L + f2) & bmask;
    L = R;
    R = t2;
  }
  return (L << d.s) | R;
}

__host__ __device__ __forceinline__ uint32_t perm_apply(uint32_t i, const PermDomain& d,
                                                        const uint32_t* keys) {
  uint32_t v = feistel_once(i, d, keys);
  while (v >= d.n) v = feistel_once(v, d, keys);
  return v;
}

}  // namespace sc
