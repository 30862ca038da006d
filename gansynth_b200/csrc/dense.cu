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

// grid: (ceil(N/256), ceil(K/kch), ceil(M/MT)), kch <= KCH; y must be zeroed (atomic split-K accumulation)
__global__ void __launch_bounds__(256) dense_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        float* __restrict__ y, int M, int K, int N, float alpha, int kch) {
  __shared__ float xs[MT][KCH];
  const int n = blockIdx.x * 256 + threadIdx.x;
  const int k0 = blockIdx.y * kch;
  const int m0 = blockIdx.z * MT;
  for (int i = threadIdx.x; i < MT * kch; i += 256) {
    int m = i / kch, k = i % kch;
    xs[m][k] = (m0 + m < M && k0 + k < K) ? x[(size_t)(m0 + m) * K + k0 + k] : 0.0f;
  }
  __syncthreads();
  if (n >= N) return;
  float acc[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m] = 0.0f;
  const int kend = (K - k0 < kch) ? K - k0 : kch;
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

// ---- vectorised forms (N % 4 == 0, M <= 8): enough 16-byte loads in flight per SM to stream the weights at HBM speed
// (the scalar kernels above kept ~4 KB per CTA in flight: 0.5 TB/s on the 8 / 17 MB matrices of the two stems).
//
// fwd: CTA = 64 column quads x 4 row lanes over a (KCHUNK x 256) block of w; the row lanes meet in shared memory, the
// K chunks in y through atomics (y zeroed by the caller).  grid (ceil(N/256), ksplit).
__global__ void __launch_bounds__(256) dense_fwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            float* __restrict__ y, int M, int K, int N, float alpha, int kchunk) {
  extern __shared__ float dsm[];
  float* xs = dsm;                         // [MT][kchunk]
  float* red = dsm + MT * kchunk;          // [4][MT][256]
  const int tn = threadIdx.x & 63, tk = threadIdx.x >> 6;
  const int n = blockIdx.x * 256 + 4 * tn;
  const int k0 = blockIdx.y * kchunk;
  const int kend = min(kchunk, K - k0);
  for (int i = threadIdx.x; i < MT * kchunk; i += 256) {
    const int m = i / kchunk, k = i - m * kchunk;
    xs[i] = (m < M && k < kend) ? x[(size_t)m * K + k0 + k] : 0.0f;
  }
  __syncthreads();
  float4 acc[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n < N) {
    const float* wp = w + (size_t)k0 * N + n;
#pragma unroll 8
    for (int k = tk; k < kend; k += 4) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(wp + (size_t)k * N));
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const float xv = xs[m * kchunk + k];
        acc[m].x = fmaf(xv, wv.x, acc[m].x); acc[m].y = fmaf(xv, wv.y, acc[m].y);
        acc[m].z = fmaf(xv, wv.z, acc[m].z); acc[m].w = fmaf(xv, wv.w, acc[m].w);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) *reinterpret_cast<float4*>(&red[(tk * MT + m) * 256 + 4 * tn]) = acc[m];
  __syncthreads();
  // thread t owns column t of the block for every m
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col < N) {
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if (m < M) {
        const float sum = red[(0 * MT + m) * 256 + threadIdx.x] + red[(1 * MT + m) * 256 + threadIdx.x] +
                          red[(2 * MT + m) * 256 + threadIdx.x] + red[(3 * MT + m) * 256 + threadIdx.x];
        atomicAdd(y + (size_t)m * N + col, alpha * sum);
      }
    }
  }
}

// dgrad: a warp owns (weight row k, segment of <= 1024 columns): 8 float4 per lane in flight, dy segment in shared
// memory, one shuffle reduction per batch row.  grid (ceil(K/8), nseg); dx zeroed by the caller when nseg > 1.
__global__ void __launch_bounds__(256) dense_dgrad_vec_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                              float* __restrict__ dx, int M, int K, int N, float alpha, int seg) {
  extern __shared__ float dsm[];           // [MT][seg]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * 8 + warp;
  const int n0 = blockIdx.y * seg;
  const int nend = min(seg, N - n0);
  for (int i = threadIdx.x; i < MT * (seg / 4); i += 256) {
    const int m = i / (seg / 4), q = i - m * (seg / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < M && 4 * q < nend) v = *reinterpret_cast<const float4*>(dy + (size_t)m * N + n0 + 4 * q);
    *reinterpret_cast<float4*>(&dsm[m * seg + 4 * q]) = v;
  }
  __syncthreads();
  if (k >= K) return;
  float acc[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m] = 0.0f;
  const float* wr = w + (size_t)k * N + n0;
#pragma unroll 8
  for (int q = lane; 4 * q < nend; q += 32) {
    const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + 4 * q));
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const float4 d = *reinterpret_cast<const float4*>(&dsm[m * seg + 4 * q]);
      acc[m] = fmaf(d.x, wv.x, fmaf(d.y, wv.y, fmaf(d.z, wv.z, fmaf(d.w, wv.w, acc[m]))));
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    const float sum = gs_warp_sum(acc[m]);
    if (lane == 0 && m < M) {
      if (gridDim.y > 1) atomicAdd(dx + (size_t)m * K + k, alpha * sum);
      else dx[(size_t)m * K + k] = alpha * sum;
    }
  }
}

// wgrad: outer products.  A thread keeps its 4 columns of dy (all batch rows) in registers and walks 8 weight rows:
// one 16-byte store (or read-modify-write when `accumulate`) per row.  grid (ceil(N/1024), ceil(K/8)), block 256.
__global__ void __launch_bounds__(256) dense_wgrad_vec_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              float* __restrict__ dw, int M, int K, int N, float alpha,
                                                              int accumulate) {
  __shared__ float xs[8][MT];
  const int n = blockIdx.x * 1024 + 4 * threadIdx.x;
  const int k0 = blockIdx.y * 8;
  if (threadIdx.x < 64) {
    const int kk = threadIdx.x >> 3, m = threadIdx.x & 7;
    xs[kk][m] = (m < M && k0 + kk < K) ? alpha * x[(size_t)m * K + k0 + kk] : 0.0f;
  }
  __syncthreads();
  if (n >= N) return;
  float4 d[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m)
    d[m] = (m < M) ? *reinterpret_cast<const float4*>(dy + (size_t)m * N + n) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    if (k0 + kk >= K) break;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const float xv = xs[kk][m];
      o.x = fmaf(xv, d[m].x, o.x); o.y = fmaf(xv, d[m].y, o.y); o.z = fmaf(xv, d[m].z, o.z); o.w = fmaf(xv, d[m].w, o.w);
    }
    float4* dst = reinterpret_cast<float4*>(dw + (size_t)(k0 + kk) * N + n);
    if (accumulate) { const float4 p = *dst; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
    *dst = o;
  }
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int gs_dense_fwd(const float* x, const float* w, float* y, int m, int k, int n, float alpha, void* stream) {
  GS_CHECK_ARG(m > 0 && k > 0 && n > 0, "dense_fwd: bad shape");
  GS_CUDA(cudaMemsetAsync(y, 0, (size_t)m * n * sizeof(float), ST));
  if (m <= MT && n % 4 == 0 && k >= 64 && ((uintptr_t)w & 15) == 0) {
    // about two CTAs per SM; K chunks of at least 16 rows, at most 512 (shared-memory copy of x)
    const int nblocks = gs_cdiv(n, 256);
    int ksplit = gs_cdiv(2 * gs_num_sms(), nblocks);
    int kchunk = gs_cdiv(k, ksplit);
    kchunk = ((kchunk < 16 ? 16 : kchunk > 512 ? 512 : kchunk) + 3) & ~3;
    ksplit = gs_cdiv(k, kchunk);
    const size_t smem = (size_t)(MT * kchunk + 4 * MT * 256) * sizeof(float);
    dense_fwd_vec_kernel<<<dim3((unsigned)nblocks, (unsigned)ksplit), 256, smem, ST>>>(x, w, y, m, k, n, alpha, kchunk);
    GS_CHECK_LAUNCH("dense_fwd_vec");
    return GS_OK;
  }
  // small matrices (the 256 -> 61 logits layer): short K chunks, so that the dependent-load chains stay short and the
  // grid still has a few dozen CTAs
  const int kch = ((size_t)k * n < ((size_t)1 << 20)) ? 8 : KCH;
  dim3 grid((unsigned)gs_cdiv(n, 256), (unsigned)gs_cdiv(k, kch), (unsigned)gs_cdiv(m, MT));
  dense_fwd_kernel<<<grid, 256, 0, ST>>>(x, w, y, m, k, n, alpha, kch);
  GS_CHECK_LAUNCH("dense_fwd");
  return GS_OK;
}
extern "C" int gs_dense_dgrad(const float* dy, const float* w, float* dx, int m, int k, int n, float alpha, void* stream) {
  GS_CHECK_ARG(m > 0 && k > 0 && n > 0, "dense_dgrad: bad shape");
  if (m <= MT && n % 4 == 0 && n >= 128 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)dy & 15) == 0) {
    const int seg = n >= 1024 ? 1024 : ((n + 127) / 128) * 128;
    const int nseg = gs_cdiv(n, seg);
    if (nseg > 1) GS_CUDA(cudaMemsetAsync(dx, 0, (size_t)m * k * sizeof(float), ST));
    dense_dgrad_vec_kernel<<<dim3((unsigned)gs_cdiv(k, 8), (unsigned)nseg), 256, (size_t)MT * seg * sizeof(float), ST>>>(
        dy, w, dx, m, k, n, alpha, seg);
    GS_CHECK_LAUNCH("dense_dgrad_vec");
    return GS_OK;
  }
  dim3 grid((unsigned)gs_cdiv(k, 4), (unsigned)gs_cdiv(m, MT));
  dense_dgrad_kernel<<<grid, 128, 0, ST>>>(dy, w, dx, m, k, n, alpha);
  GS_CHECK_LAUNCH("dense_dgrad");
  return GS_OK;
}
extern "C" int gs_dense_wgrad(const float* x, const float* dy, float* dw, int m, int k, int n, float alpha, void* stream) {
  GS_CHECK_ARG(m > 0 && k > 0 && n > 0, "dense_wgrad: bad shape");
  if (m <= MT && n % 4 == 0 && ((uintptr_t)dw & 15) == 0 && ((uintptr_t)dy & 15) == 0) {
    dense_wgrad_vec_kernel<<<dim3((unsigned)gs_cdiv(n, 1024), (unsigned)gs_cdiv(k, 8)), 256, 0, ST>>>(x, dy, dw, m, k, n, alpha, 0);
    GS_CHECK_LAUNCH("dense_wgrad_vec");
    return GS_OK;
  }
  size_t total = (size_t)k * n;
  size_t b = (total + 255) / 256;
  size_t cap = (size_t)gs_num_sms() * 32;
  dense_wgrad_kernel<<<(unsigned)(b < cap ? b : cap), 256, 0, ST>>>(x, dy, dw, m, k, n, alpha);
  GS_CHECK_LAUNCH("dense_wgrad");
  return GS_OK;
}
