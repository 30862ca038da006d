// 1x1 convolutions with a tiny channel count on one side: the 2-channel image side of the colour blocks
// (reference networks.py:95-107 to-RGB, :230-242 from-RGB) and their gradients.  These are HBM-bound
// per-pixel maps (expand 2 -> C, reduce C -> 2) and a pixel reduction (filter gradient); the generic
// tiled kernels would waste > 90 % of their tile on them.
#pragma once
#include "common.cuh"

// y[p][n] = act(alpha * sum_{k<KD} x[p][k] * B(k,n) + bias[n]);  ndim % 4 == 0, ndim <= 1024.
// B(k,n) = w[k*ndim+n] (w_is_kn) or w[n*KD+k].
template <int KD>
__global__ void __launch_bounds__(256) conv1x1_expand_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ y,
                                                             long long npix, int ndim, int w_is_kn, float alpha, int act) {
  extern __shared__ float sB[];   // [KD][ndim] then bias [ndim]
  for (int i = threadIdx.x; i < KD * ndim; i += blockDim.x) {
    int k = i / ndim, n = i % ndim;
    sB[i] = alpha * (w_is_kn ? w[(size_t)k * ndim + n] : w[(size_t)n * KD + k]);
  }
  for (int i = threadIdx.x; i < ndim; i += blockDim.x) sB[KD * ndim + i] = bias ? bias[i] : 0.0f;
  __syncthreads();
  const int n4 = ndim / 4;
  const long long total = npix * n4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long p = idx / n4;
    const int n = (int)(idx % n4) * 4;
    float4 o = *reinterpret_cast<const float4*>(sB + KD * ndim + n);
#pragma unroll
    for (int k = 0; k < KD; ++k) {
      const float xv = __ldg(x + p * KD + k);
      const float4 b = *reinterpret_cast<const float4*>(sB + k * ndim + n);
      o.x = fmaf(xv, b.x, o.x); o.y = fmaf(xv, b.y, o.y); o.z = fmaf(xv, b.z, o.z); o.w = fmaf(xv, b.w, o.w);
    }
    if (act == 1) { o.x = gs_lrelu(o.x); o.y = gs_lrelu(o.y); o.z = gs_lrelu(o.z); o.w = gs_lrelu(o.w); }
    *reinterpret_cast<float4*>(y + p * ndim + n) = o;
  }
}

// y[p][n] for n < ND <= 4, kdim % 4 == 0: LPP = min(32, kdim/4) lanes cooperate on one pixel.
template <int ND>
__global__ void __launch_bounds__(256) conv1x1_reduce_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ y,
                                                             long long npix, int kdim, int w_is_kn, float alpha, int act) {
  extern __shared__ float sB[];   // [ND][kdim]
  for (int i = threadIdx.x; i < ND * kdim; i += blockDim.x) {
    int n = i / kdim, k = i % kdim;
    sB[i] = alpha * (w_is_kn ? w[(size_t)k * ND + n] : w[(size_t)n * kdim + k]);
  }
  __syncthreads();
  int lpp = kdim / 4;
  if (lpp > 32) lpp = 32;
  const int ppw = 32 / lpp;                         // pixels per warp instruction
  const int lane = threadIdx.x & 31, sub = lane % lpp, slot = lane / lpp;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  // U pixel groups per iteration: all their loads are issued before the first reduction (memory-level parallelism)
  constexpr int U = 8;
  const bool one_pass = kdim <= lpp * 4;      // every lane holds one float4 of its pixel: all U loads first, then the math
  for (long long p0 = warp * ppw * U; p0 < npix; p0 += nwarps * ppw * U) {
    float acc[U][ND];
    if (one_pass) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = p0 + (long long)u * ppw + slot;
        v[u] = p < npix ? __ldg(reinterpret_cast<const float4*>(x + p * kdim + sub * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int n = 0; n < ND; ++n) {
        const float4 b = *reinterpret_cast<const float4*>(sB + n * kdim + sub * 4);
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u][n] = v[u].x * b.x + v[u].y * b.y + v[u].z * b.z + v[u].w * b.w;
      }
    } else
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p0 + (long long)u * ppw + slot;
#pragma unroll
      for (int n = 0; n < ND; ++n) acc[u][n] = 0.0f;
      if (p < npix) {
        for (int k = sub * 4; k < kdim; k += lpp * 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(x + p * kdim + k));
#pragma unroll
          for (int n = 0; n < ND; ++n) {
            const float4 b = *reinterpret_cast<const float4*>(sB + n * kdim + k);
            acc[u][n] += v.x * b.x + v.y * b.y + v.z * b.z + v.w * b.w;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p0 + (long long)u * ppw + slot;
#pragma unroll
      for (int n = 0; n < ND; ++n)
        for (int o = lpp >> 1; o > 0; o >>= 1) acc[u][n] += __shfl_xor_sync(0xffffffffu, acc[u][n], o);
      if (sub == 0 && p < npix) {
#pragma unroll
        for (int n = 0; n < ND; ++n) {
          float v = acc[u][n] + (bias ? bias[n] : 0.0f);
          y[p * ND + n] = act == 1 ? gs_lrelu(v) : v;
        }
      }
    }
  }
}

// dw[c*sc + j*sj] += alpha * sum_p wide[p][c] * narrow[p][j],  j < ND <= 4, wide channels W <= 256, W % 4 == 0.
// A thread owns one float4 of wide channels and walks pixels, four of them in flight (16-byte loads); the block meets in
// shared memory, one atomic per (block, output element).
template <int ND>
__global__ void __launch_bounds__(256) conv1x1_w_kernel(const float* __restrict__ wide, const float* __restrict__ narrow,
                                                        float* __restrict__ dw, long long npix, int W, int sc, int sj,
                                                        float alpha, long long pix_per_block) {
  __shared__ float red[256 * 4 * ND];
  const int quads = W / 4;                     // threads per pixel
  const int lanes = 256 / quads;               // pixel lanes per block (W in {32, 64, 128, 256})
  const int q = threadIdx.x % quads, pl = threadIdx.x / quads;
  const long long p0 = blockIdx.x * pix_per_block;
  const long long p1 = p0 + pix_per_block < npix ? p0 + pix_per_block : npix;
  float4 acc[ND];
#pragma unroll
  for (int j = 0; j < ND; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int U = 4;
  for (long long p = p0 + pl; p < p1; p += (long long)lanes * U) {
    float4 v[U];
    float nv[U][ND];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long pp = p + (long long)u * lanes;
      const bool ok = pp < p1;
      v[u] = ok ? __ldg(reinterpret_cast<const float4*>(wide + pp * W + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < ND; ++j) nv[u][j] = ok ? __ldg(narrow + pp * ND + j) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < ND; ++j) {
        acc[j].x = fmaf(v[u].x, nv[u][j], acc[j].x); acc[j].y = fmaf(v[u].y, nv[u][j], acc[j].y);
        acc[j].z = fmaf(v[u].z, nv[u][j], acc[j].z); acc[j].w = fmaf(v[u].w, nv[u][j], acc[j].w);
      }
  }
#pragma unroll
  for (int j = 0; j < ND; ++j) *reinterpret_cast<float4*>(&red[(j * 256 + threadIdx.x) * 4]) = acc[j];
  __syncthreads();
  if (threadIdx.x < W) {
    const int c = threadIdx.x, cq = c >> 2, ce = c & 3;
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      float s = 0.0f;
      for (int l = 0; l < lanes; ++l) s += red[(j * 256 + l * quads + cq) * 4 + ce];
      atomicAdd(dw + (size_t)c * sc + (size_t)j * sj, alpha * s);
    }
  }
}
