// Error reporting + launch accounting for libsdxl_b200.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace b2 {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace b2

extern "C" int b2_version(void) { return 100; }
extern "C" const char* b2_last_error(void) { return b2::g_err; }
extern "C" long long b2_launch_count(void) { return b2::g_launches.load(); }
