#include <stdarg.h>

#include "common.cuh"

namespace dexnerf {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace dexnerf

extern "C" DEXNERF_API int dexnerf_abi_version(void) { return DEXNERF_ABI_VERSION; }
extern "C" DEXNERF_API const char* dexnerf_last_error(void) { return dexnerf::g_err; }
