// Error reporting and version for libgansynth_b200.so (see include/gansynth_b200.h).
#include <stdarg.h>
#include "common.cuh"
#include "gansynth_b200.h"

static thread_local char g_err[512] = "";

void gs_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* gs_last_error(void) { return g_err; }
extern "C" int gs_version(void) { return 100; }
