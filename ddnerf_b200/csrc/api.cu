// Library-level entry points: version, thread-local error string, device probe.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace ddnerf {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace ddnerf

extern "C" DDNERF_EXPORT int ddnerf_version(void) { return DDNERF_ABI_VERSION; }
extern "C" DDNERF_EXPORT const char* ddnerf_last_error(void) { return ddnerf::g_err; }
extern "C" DDNERF_EXPORT int ddnerf_device_is_sm100(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10;
}
extern "C" DDNERF_EXPORT int64_t ddnerf_launch_count(void) { return (int64_t)ddnerf::g_launches.load(); }
