// Error reporting and version entry points of the C ABI (include/fsnet_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include "common.cuh"

namespace fsnet {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("FSNET_PDL"); on = e ? atoi(e) != 0 : 1; }
  return on != 0;
}
}  // namespace fsnet

extern "C" int fsnet_abi_version(void) { return 1; }
extern "C" const char* fsnet_last_error(void) { return fsnet::g_err; }
