// Error channel and device queries shared by all translation units of libsc_b200.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace sc {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
      cached_sms = sms;
    cached_dev = dev;
  }
  return cached_sms;
}

}  // namespace sc

extern "C" int sc_version(void) { return 100; }
extern "C" const char* sc_last_error(void) { return sc::g_error; }
namespace sc { long long launches(); }
extern "C" long long sc_launch_count(void) { return sc::launches(); }

extern "C" int sc_philox_permutation_host(uint64_t seed, int64_t perm_index, int64_t n,
                                          int32_t* out_host) {
  SC_CHECK_ARG(out_host && n >= 1 && n < (1ll << 31), "sc_philox_permutation_host: bad argument");
  uint32_t keys[sc::kFeistelRounds];
  sc::perm_round_keys(seed, (uint64_t)perm_index, keys);
  sc::PermDomain d = sc::make_perm_domain((uint32_t)n);
  for (int64_t i = 0; i < n; ++i) out_host[i] = (int32_t)sc::perm_apply((uint32_t)i, d, keys);
  return SC_OK;
}
