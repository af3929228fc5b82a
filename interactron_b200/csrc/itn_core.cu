// Error reporting, version, launch accounting for the C ABI.
#include <cstring>

#include "itn_common.cuh"

namespace itn {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace itn

extern "C" const char* itn_last_error(void) { return itn::g_err; }
extern "C" const char* itn_version(void) { return "interactron_b200 0.1 sm_100a"; }
extern "C" long long itn_launch_count(void) {
  return itn::g_launches.load(std::memory_order_relaxed);
}
