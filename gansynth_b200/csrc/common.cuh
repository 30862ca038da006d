// Shared helpers for the gansynth_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define GS_OK 0
#define GS_ERR_ARG -1
#define GS_ERR_CUDA -2
#define GS_ERR_UNSUPPORTED -3

void gs_set_error(const char* fmt, ...);

#define GS_CHECK_ARG(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      gs_set_error(__VA_ARGS__);           \
      return GS_ERR_ARG;                   \
    }                                      \
  } while (0)

#define GS_CHECK_LAUNCH(name)                                                      \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      gs_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));        \
      return GS_ERR_CUDA;                                                          \
    }                                                                              \
  } while (0)

#define GS_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      gs_set_error("%s failed: %s", #call, cudaGetErrorString(e__));               \
      return GS_ERR_CUDA;                                                          \
    }                                                                              \
  } while (0)

// ---- library context (include/gansynth_b200.h: gs_context_create / gs_context_bind) ---------------------------
// All state the library keeps between calls lives here, in caller-owned memory: the device workspace the caller passed
// (spectral twiddle tables | scratch slot for split weights used once | cache of split PARAMETER weights) and the host
// table describing what the cache holds.  One context is bound per host thread; a context serves one stream at a time.
struct GsPrepKey {
  const float* w;
  int kdim, ndim, nt, kc, kn, flip;
  size_t off;
};
struct gs_context {
  unsigned char* ws;
  size_t bytes;
  size_t tables_off, scratch_off, scratch, cache_off, cache, used;
  GsPrepKey prep[512];
  int nprep;
  int cache_full_warned;
  bool tables_ready;
};
constexpr size_t GS_WS_TABLES = 64 << 10;            // twiddle tables (16 KB used)
constexpr size_t GS_WS_SCRATCH = (size_t)8 << 20;    // largest split weight: 9 * 256 * 256 * 2 * 2 bytes = 2.4 MB
constexpr size_t GS_WS_CACHE = (size_t)256 << 20;
gs_context* gs_bound_context();                      // context of the calling thread, or nullptr
#define GS_NEED_CONTEXT(ctx, what)                                                                         \
  gs_context* ctx = gs_bound_context();                                                                    \
  GS_CHECK_ARG(ctx != nullptr, what ": no context bound to this thread (gs_context_create + gs_context_bind)")

static inline int gs_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

static inline int gs_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float gs_lrelu(float v) { return v > 0.0f ? v : 0.2f * v; }
__device__ __forceinline__ float gs_lrelu_slope(float y) { return y > 0.0f ? 1.0f : 0.2f; }

__device__ __forceinline__ float gs_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
