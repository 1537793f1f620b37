// Error reporting and version entry points of the C ABI (include/fsnet_b200.h).
#include <stdarg.h>
#include "common.cuh"

namespace fsnet {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace fsnet

extern "C" int fsnet_abi_version(void) { return 1; }
extern "C" const char* fsnet_last_error(void) { return fsnet::g_err; }
