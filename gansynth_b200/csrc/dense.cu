// Small-batch dense layers (reference ops.py:183-201 tf.matmul call site and its gradients).
// The batch (M) is 8 on the training path, so all three forms are weight-bandwidth bound GEMVs:
//   fwd   y[M,N]  = alpha * x[M,K]  @ w[K,N]
//   dgrad dx[M,K] = alpha * dy[M,N] @ w[K,N]^T
//   wgrad dw[K,N] = alpha * x[M,K]^T @ dy[M,N]
#include "common.cuh"
#include "gansynth_b200.h"

namespace {

constexpr int MT = 8;     // batch rows per CTA
constexpr int KCH = 64;   // split-K chunk of the forward kernel

// grid: (ceil(N/256), ceil(K/KCH), ceil(M/MT)); y must be zeroed (atomic split-K accumulation)
__global__ void __launch_bounds__(256) dense_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        float* __restrict__ y, int M, int K, int N, float alpha) {
  __shared__ float xs[MT][KCH];
  const int n = blockIdx.x * 256 + threadIdx.x;
  const int k0 = blockIdx.y * KCH;
  const int m0 = blockIdx.z * MT;
  for (int i = threadIdx.x; i < MT * KCH; i += 256) {
    int m = i / KCH, k = i % KCH;
    xs[m][k] = (m0 + m < M && k0 + k < K) ? x[(size_t)(m0 + m) * K + k0 + k] : 0.0f;
  }
  __syncthreads();
  if (n >= N) return;
  float acc[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m] = 0.0f;
  const int kend = (K - k0 < KCH) ? K - k0 : KCH;
#pragma unroll 4
  for (int k = 0; k < kend; ++k) {
    float wv = w[(size_t)(k0 + k) * N + n];
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[m] = fmaf(xs[m][k], wv, acc[m]);
  }
#pragma unroll
  for (int m = 0; m < MT; ++m)
    if (m0 + m < M) atomicAdd(y + (size_t)(m0 + m) * N + n, alpha * acc[m]);
}

// grid: (ceil(K/4), ceil(M/MT)); block 128 = 4 warps, one weight row per warp
constexpr int NCH = 1024;
__global__ void __launch_bounds__(128) dense_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                          float* __restrict__ dx, int M, int K, int N, float alpha) {
  __shared__ float ds[MT][NCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * 4 + warp;
  const int m0 = blockIdx.y * MT;
  float acc[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m] = 0.0f;
  for (int n0 = 0; n0 < N; n0 += NCH) {
    __syncthreads();
    for (int i = threadIdx.x; i < MT * NCH; i += 128) {
      int m = i / NCH, n = i % NCH;
      ds[m][n] = (m0 + m < M && n0 + n < N) ? dy[(size_t)(m0 + m) * N + n0 + n] : 0.0f;
    }
    __syncthreads();
    if (k < K) {
      const int nend = (N - n0 < NCH) ? N - n0 : NCH;
      for (int n = lane; n < nend; n += 32) {
        float wv = w[(size_t)k * N + n0 + n];
#pragma unroll
        for (int m = 0; m < MT; ++m) acc[m] = fmaf(ds[m][n], wv, acc[m]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    float s = gs_warp_sum(acc[m]);
    if (lane == 0 && k < K && m0 + m < M) dx[(size_t)(m0 + m) * K + k] = alpha * s;
  }
}

__global__ void dense_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int M,
                                   int K, int N, float alpha) {
  size_t total = (size_t)K * N;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int n = (int)(i % N);
    int k = (int)(i / N);
    float s = 0.0f;
    for (int m = 0; m < M; ++m) s = fmaf(x[(size_t)m * K + k], dy[(size_t)m * N + n], s);
    dw[i] = alpha * s;
  }
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int gs_dense_fwd(const float* x, const float* w, float* y, int m, int k, int n, float alpha, void* stream) {
  GS_CHECK_ARG(m > 0 && k > 0 && n > 0, "dense_fwd: bad shape");
  GS_CUDA(cudaMemsetAsync(y, 0, (size_t)m * n * sizeof(float), ST));
  dim3 grid((unsigned)gs_cdiv(n, 256), (unsigned)gs_cdiv(k, KCH), (unsigned)gs_cdiv(m, MT));
  dense_fwd_kernel<<<grid, 256, 0, ST>>>(x, w, y, m, k, n, alpha);
  GS_CHECK_LAUNCH("dense_fwd");
  return GS_OK;
}
extern "C" int gs_dense_dgrad(const float* dy, const float* w, float* dx, int m, int k, int n, float alpha, void* stream) {
  GS_CHECK_ARG(m > 0 && k > 0 && n > 0, "dense_dgrad: bad shape");
  dim3 grid((unsigned)gs_cdiv(k, 4), (unsigned)gs_cdiv(m, MT));
  dense_dgrad_kernel<<<grid, 128, 0, ST>>>(dy, w, dx, m, k, n, alpha);
  GS_CHECK_LAUNCH("dense_dgrad");
  return GS_OK;
}
extern "C" int gs_dense_wgrad(const float* x, const float* dy, float* dw, int m, int k, int n, float alpha, void* stream) {
  GS_CHECK_ARG(m > 0 && k > 0 && n > 0, "dense_wgrad: bad shape");
  size_t total = (size_t)k * n;
  size_t b = (total + 255) / 256;
  size_t cap = (size_t)gs_num_sms() * 32;
  dense_wgrad_kernel<<<(unsigned)(b < cap ? b : cap), 256, 0, ST>>>(x, dy, dw, m, k, n, alpha);
  GS_CHECK_LAUNCH("dense_wgrad");
  return GS_OK;
}
