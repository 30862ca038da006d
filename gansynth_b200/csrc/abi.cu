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

// ---- context ---------------------------------------------------------------------------------------------------
static thread_local gs_context* g_ctx = nullptr;
gs_context* gs_bound_context() { return g_ctx; }

extern "C" size_t gs_workspace_bytes(void) { return GS_WS_TABLES + GS_WS_SCRATCH + GS_WS_CACHE; }
extern "C" size_t gs_workspace_min_bytes(void) { return GS_WS_TABLES + GS_WS_SCRATCH; }

extern "C" int gs_context_create(void* workspace, size_t bytes, gs_context** out) {
  GS_CHECK_ARG(out != nullptr, "context_create: out is null");
  GS_CHECK_ARG(workspace != nullptr && ((uintptr_t)workspace & 255) == 0, "context_create: workspace must be a 256-byte aligned device pointer");
  GS_CHECK_ARG(bytes >= GS_WS_TABLES + GS_WS_SCRATCH, "context_create: workspace of %zu bytes is below gs_workspace_min_bytes() = %zu",
               bytes, GS_WS_TABLES + GS_WS_SCRATCH);
  gs_context* c = new gs_context();
  c->ws = static_cast<unsigned char*>(workspace);
  c->bytes = bytes;
  c->tables_off = 0;
  c->scratch_off = GS_WS_TABLES;
  c->scratch = GS_WS_SCRATCH;
  c->cache_off = GS_WS_TABLES + GS_WS_SCRATCH;
  c->cache = bytes - c->cache_off;
  c->used = 0;
  c->nprep = 0;
  c->cache_full_warned = 0;
  c->tables_ready = false;
  *out = c;
  return GS_OK;
}
extern "C" int gs_context_destroy(gs_context* ctx) {
  if (ctx == g_ctx) g_ctx = nullptr;
  delete ctx;
  return GS_OK;
}
extern "C" int gs_context_bind(gs_context* ctx) {
  g_ctx = ctx;
  return GS_OK;
}
