// Error reporting / bookkeeping entry points of the C ABI (include/tomoenc.h).
#include "te_common.cuh"

namespace te {
static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace te

extern "C" int te_version(void) { return 100; }
extern "C" const char* te_last_error(void) { return te::g_err; }
extern "C" int64_t te_launch_count(void) { return te::g_launches.load(); }
extern "C" int64_t te_launch_count_add(int64_t n) {
  te::count_launch(static_cast<int>(n));                  // n < 0: a capture pass went through the counting entry points
  return te::g_launches.load();
}
