// api.cu -- version + thread-local error message for the C ABI (include/vqb.h).
#include <stdarg.h>

#include "common.cuh"

namespace vqb {
int64_t g_launch_count = 0;
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace vqb

extern "C" int vqb_version(void) { return VQB_VERSION; }
extern "C" const char* vqb_last_error(void) { return vqb::g_err; }
extern "C" int64_t vqb_launch_count(void) { return vqb::g_launch_count; }
