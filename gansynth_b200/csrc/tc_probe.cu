// tcgen05 self-test ("probe"): one CTA, one 128 x N accumulator, operands staged in shared memory by
// ordinary threads in the SWIZZLE_NONE core-matrix layout the implicit-GEMM convolution uses:
//   mode 0 (K-major A and B, conv forward / dgrad form):
//       D[m][n] = sum_k bf16(A[pix(m)][k]) * bf16(B[n][k]),  pix(m) = (m/8)*gstride + m%8 + shift
//   mode 1 (MN-major A and B, filter-gradient form):
//       D[m][n] = sum_j bf16(A[pixk(j)][m]) * bf16(B[pixk(j)][n]),  pixk(j) = (j/8)*gstride + j%8 + shift
// `shift` / `gstride` exercise exactly the "shifted window into a halo tile" addressing (descriptor start
// address and stride fields) that gives the 3x3 taps without re-staging the tile.
#include "common.cuh"
#include "tc_probe.h"
#include "tc_common.cuh"

namespace {

__global__ void __launch_bounds__(160) tc_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                       float* __restrict__ D, int K, int N, int rows_a, int rows_b,
                                                       int shift, int gstride, int mode, int reps,
                                                       long long* __restrict__ cycles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // A region then B region, each [chunk q][row p] of 16-byte vectors (8 bf16)
  const int a_chunks = (mode == 0) ? K / 8 : 128 / 8;
  const int b_chunks = (mode == 0) ? K / 8 : N / 8;
  uint4* As = reinterpret_cast<uint4*>(smem_raw);
  uint4* Bs = As + a_chunks * rows_a;
  const int a_cols = (mode == 0) ? K : 128;
  const int b_cols = (mode == 0) ? K : N;
  for (int i = tid; i < a_chunks * rows_a; i += blockDim.x) {
    int q = i / rows_a, p = i % rows_a;
    const float* src = A + (size_t)p * a_cols + q * 8;
    uint4 v;
    v.x = tc::pack_bf16(__float2bfloat16_rn(src[0]), __float2bfloat16_rn(src[1]));
    v.y = tc::pack_bf16(__float2bfloat16_rn(src[2]), __float2bfloat16_rn(src[3]));
    v.z = tc::pack_bf16(__float2bfloat16_rn(src[4]), __float2bfloat16_rn(src[5]));
    v.w = tc::pack_bf16(__float2bfloat16_rn(src[6]), __float2bfloat16_rn(src[7]));
    As[q * rows_a + p] = v;
  }
  for (int i = tid; i < b_chunks * rows_b; i += blockDim.x) {
    int q = i / rows_b, p = i % rows_b;
    const float* src = B + (size_t)p * b_cols + q * 8;
    uint4 v;
    v.x = tc::pack_bf16(__float2bfloat16_rn(src[0]), __float2bfloat16_rn(src[1]));
    v.y = tc::pack_bf16(__float2bfloat16_rn(src[2]), __float2bfloat16_rn(src[3]));
    v.z = tc::pack_bf16(__float2bfloat16_rn(src[4]), __float2bfloat16_rn(src[5]));
    v.w = tc::pack_bf16(__float2bfloat16_rn(src[6]), __float2bfloat16_rn(src[7]));
    Bs[q * rows_b + p] = v;
  }
  tc::fence_proxy_async();
  int ncols = 32;
  while (ncols < N) ncols <<= 1;
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_fence_init();
  }
  if (warp == 4) tc::tmem_alloc(&tmem_base_s, (uint32_t)ncols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 4) {
    // descriptors of the (at most 8) K slices are built once: the timed loop is UTCHMMA instructions only
    const uint32_t a0 = tc::smem_u32(As), b0 = tc::smem_u32(Bs);
    const uint32_t plane_a = (uint32_t)rows_a * 16u, plane_b = (uint32_t)rows_b * 16u;
    const uint32_t idesc = tc::idesc_bf16_f32(N, mode, mode);
    uint64_t da[8], db[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int ss = (s < K / 16) ? s : 0;
      if (mode == 0) {
        da[s] = tc::smem_desc(a0 + 2u * ss * plane_a + (uint32_t)shift * 16u, plane_a, (uint32_t)gstride * 16u);
        db[s] = tc::smem_desc(b0 + 2u * ss * plane_b, plane_b, 128u);
      } else {
        const uint32_t poff = (uint32_t)(ss * 2 * gstride + shift) * 16u;
        da[s] = tc::smem_desc(a0 + poff, (uint32_t)gstride * 16u, plane_a);
        db[s] = tc::smem_desc(b0 + poff, (uint32_t)gstride * 16u, plane_b);
      }
    }
    const int nsl = K / 16;
    const long long t_start = clock64();
    if (tc::elect_one()) {
      for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
          if (s < nsl) tc::mma_bf16(tmem_base, da[s], db[s], idesc, (s > 0 || rep > 0) ? 1u : 0u);
      }
      tc::mma_commit(&bar);
    }
    __syncwarp();
    tc::mbar_wait(&bar, 0);
    if (cycles && lane == 0) *cycles = clock64() - t_start;
  }
  if (warp < 4) {
    tc::mbar_wait(&bar, 0);
    tc::tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tc::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j < N) D[(size_t)row * N + c0 + j] = v[j];
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem_base, (uint32_t)ncols);
}

}  // namespace

static int tc_probe_impl(const float* a, const float* b, float* d, int k, int n, int rows_a, int rows_b, int shift,
                         int gstride, int mode, int reps, long long* cycles, void* stream) {
  GS_CHECK_ARG(mode == 0 || mode == 1, "tc_probe: mode must be 0 or 1");
  GS_CHECK_ARG(k > 0 && k % 16 == 0 && n >= 16 && n <= 256 && n % 16 == 0, "tc_probe: need K %% 16 == 0, 16 <= N <= 256, N %% 16 == 0");
  GS_CHECK_ARG(gstride >= 8 && shift >= 0, "tc_probe: gstride >= 8, shift >= 0");
  size_t a_bytes, b_bytes;
  if (mode == 0) {
    GS_CHECK_ARG(rows_a >= 15 * gstride + 8 + shift && rows_b == n, "tc_probe: mode 0 needs rows_a >= 15*gstride+8+shift, rows_b == N");
    a_bytes = (size_t)(k / 8) * rows_a * 16;
    b_bytes = (size_t)(k / 8) * rows_b * 16;
  } else {
    int need = (k / 8 - 1) * gstride + 8 + shift;
    GS_CHECK_ARG(rows_a >= need && rows_b == rows_a, "tc_probe: mode 1 needs rows_a == rows_b >= (K/8-1)*gstride+8+shift");
    a_bytes = (size_t)16 * rows_a * 16;
    b_bytes = (size_t)(n / 8) * rows_b * 16;
  }
  size_t smem = a_bytes + b_bytes;
  GS_CHECK_ARG(smem <= 200 * 1024, "tc_probe: operands need %zu bytes of shared memory", smem);
  GS_CUDA(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_probe_kernel<<<1, 160, smem, (cudaStream_t)stream>>>(a, b, d, k, n, rows_a, rows_b, shift, gstride, mode, reps, cycles);
  GS_CHECK_LAUNCH("tc_probe");
  return GS_OK;
}

extern "C" int gs_tc_probe(const float* a, const float* b, float* d, int k, int n, int rows_a, int rows_b, int shift,
                           int gstride, int mode, void* stream) {
  return tc_probe_impl(a, b, d, k, n, rows_a, rows_b, shift, gstride, mode, 1, nullptr, stream);
}
// timing variant: the MMA sequence is issued `reps` times back to back by the one issuing thread;
// *cycles (device) receives the SM clock ticks from first issue to completion
extern "C" int gs_tc_probe_time(const float* a, const float* b, float* d, int k, int n, int rows_a, int rows_b,
                                int shift, int gstride, int mode, int reps, long long* cycles, void* stream) {
  GS_CHECK_ARG(reps >= 1 && cycles != nullptr, "tc_probe_time: reps >= 1 and a cycles pointer are required");
  return tc_probe_impl(a, b, d, k, n, rows_a, rows_b, shift, gstride, mode, reps, cycles, stream);
}
