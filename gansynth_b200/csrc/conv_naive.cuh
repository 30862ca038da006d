// Naive NHWC convolution kernels: one thread per output element, any channel count, ksize 1|3,
// stride 1|2 (TF SAME: pad_before = (ksize==3 && stride==1) ? 1 : 0).
// They are the on-device correctness anchor for the tiled / tensor-core kernels and the path for
// odd channel counts (2-channel image side of the colour blocks, the 257-channel stddev conv).
#pragma once
#include "common.cuh"

struct ConvGeom {
  int n, h, w, ci, co;  // h,w = spatial size of the LARGE side (conv input / dgrad output)
  int oh, ow;           // small side: h/stride, w/stride
  int ksize, stride, pb;
  int wswap;            // 0: w[kh][kw][ci][co]   1: w[kh][kw][co][ci]
  float alpha;
  int act;              // 0 none, 1 leaky-relu(0.2)
};

__device__ __forceinline__ float gs_wt(const float* __restrict__ w, const ConvGeom& g, int tap, int ci, int co) {
  return g.wswap ? w[((size_t)tap * g.co + co) * g.ci + ci] : w[((size_t)tap * g.ci + ci) * g.co + co];
}

// y[n,oh,ow,co] = alpha * sum x[n, oh*s+kh-pb, ow*s+kw-pb, ci] * Wt(kh,kw,ci,co) (+bias[co]) (act)
__global__ void conv_c_naive_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                    const float* __restrict__ bias, float* __restrict__ y, ConvGeom g) {
  size_t total = (size_t)g.n * g.oh * g.ow * g.co;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int co = (int)(idx % g.co);
    size_t p = idx / g.co;
    int ow = (int)(p % g.ow);
    p /= g.ow;
    int oh = (int)(p % g.oh);
    int n = (int)(p / g.oh);
    float acc = 0.0f;
    for (int kh = 0; kh < g.ksize; ++kh) {
      int ih = oh * g.stride + kh - g.pb;
      if (ih < 0 || ih >= g.h) continue;
      for (int kw = 0; kw < g.ksize; ++kw) {
        int iw = ow * g.stride + kw - g.pb;
        if (iw < 0 || iw >= g.w) continue;
        const float* xp = x + (((size_t)n * g.h + ih) * g.w + iw) * g.ci;
        int tap = kh * g.ksize + kw;
        for (int ci = 0; ci < g.ci; ++ci) acc = fmaf(xp[ci], gs_wt(w, g, tap, ci, co), acc);
      }
    }
    acc *= g.alpha;
    if (bias) acc += bias[co];
    if (g.act == 1) acc = gs_lrelu(acc);
    y[idx] = acc;
  }
}

// dx[n,ih,iw,ci] = alpha * sum_{kh,kw,co : (ih+pb-kh) % s == 0 ...} dy[n,(ih+pb-kh)/s,(iw+pb-kw)/s,co] * Wt(kh,kw,ci,co)
__global__ void conv_t_naive_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                    const float* __restrict__ bias, float* __restrict__ dx, ConvGeom g) {
  size_t total = (size_t)g.n * g.h * g.w * g.ci;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int ci = (int)(idx % g.ci);
    size_t p = idx / g.ci;
    int iw = (int)(p % g.w);
    p /= g.w;
    int ih = (int)(p % g.h);
    int n = (int)(p / g.h);
    float acc = 0.0f;
    for (int kh = 0; kh < g.ksize; ++kh) {
      int th = ih + g.pb - kh;
      if (th < 0 || (th % g.stride) != 0) continue;
      int oh = th / g.stride;
      if (oh >= g.oh) continue;
      for (int kw = 0; kw < g.ksize; ++kw) {
        int tw = iw + g.pb - kw;
        if (tw < 0 || (tw % g.stride) != 0) continue;
        int ow = tw / g.stride;
        if (ow >= g.ow) continue;
        const float* yp = dy + (((size_t)n * g.oh + oh) * g.ow + ow) * g.co;
        int tap = kh * g.ksize + kw;
        for (int co = 0; co < g.co; ++co) acc = fmaf(yp[co], gs_wt(w, g, tap, ci, co), acc);
      }
    }
    acc *= g.alpha;
    if (bias) acc += bias[ci];
    if (g.act == 1) acc = gs_lrelu(acc);
    dx[idx] = acc;
  }
}

// Same contract, one WARP per output element (lanes stride over the contraction channel co): for outputs
// with few elements and a long contraction (the 1-channel stddev branch) the thread-per-output kernel is
// a latency chain of thousands of dependent loads.
__global__ void conv_t_naive_warp_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                         const float* __restrict__ bias, float* __restrict__ dx, ConvGeom g) {
  const size_t total = (size_t)g.n * g.h * g.w * g.ci;
  const int lane = threadIdx.x & 31;
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t idx = warp; idx < total; idx += nwarps) {
    int ci = (int)(idx % g.ci);
    size_t p = idx / g.ci;
    int iw = (int)(p % g.w);
    p /= g.w;
    int ih = (int)(p % g.h);
    int n = (int)(p / g.h);
    float acc = 0.0f;
    for (int kh = 0; kh < g.ksize; ++kh) {
      int th = ih + g.pb - kh;
      if (th < 0 || (th % g.stride) != 0) continue;
      int oh = th / g.stride;
      if (oh >= g.oh) continue;
      for (int kw = 0; kw < g.ksize; ++kw) {
        int tw = iw + g.pb - kw;
        if (tw < 0 || (tw % g.stride) != 0) continue;
        int ow = tw / g.stride;
        if (ow >= g.ow) continue;
        const float* yp = dy + (((size_t)n * g.oh + oh) * g.ow + ow) * g.co;
        int tap = kh * g.ksize + kw;
        for (int co = lane; co < g.co; co += 32) acc = fmaf(yp[co], gs_wt(w, g, tap, ci, co), acc);
      }
    }
    acc = gs_warp_sum(acc);
    if (lane == 0) {
      acc *= g.alpha;
      if (bias) acc += bias[ci];
      if (g.act == 1) acc = gs_lrelu(acc);
      dx[idx] = acc;
    }
  }
}

// dw(kh,kw,ci,co) += alpha * sum over a pixel chunk of x[n,oh*s+kh-pb,ow*s+kw-pb,ci] * dy[n,oh,ow,co]
// grid.x covers (tap,ci,co) in blocks of 128 threads, grid.y = pixel chunks; dw must be zeroed first.
__global__ void conv_w_naive_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                    float* __restrict__ dw, ConvGeom g, int pix_per_chunk) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  int nel = g.ksize * g.ksize * g.ci * g.co;
  if (e >= nel) return;
  int tap, ci, co;
  if (g.wswap) {
    ci = e % g.ci;
    co = (e / g.ci) % g.co;
    tap = e / (g.ci * g.co);
  } else {
    co = e % g.co;
    ci = (e / g.co) % g.ci;
    tap = e / (g.ci * g.co);
  }
  int kh = tap / g.ksize, kw = tap % g.ksize;
  long long npix = (long long)g.n * g.oh * g.ow;
  long long p0 = (long long)blockIdx.y * pix_per_chunk;
  long long p1 = p0 + pix_per_chunk < npix ? p0 + pix_per_chunk : npix;
  float acc = 0.0f;
  for (long long p = p0; p < p1; ++p) {
    int ow = (int)(p % g.ow);
    long long q = p / g.ow;
    int oh = (int)(q % g.oh);
    int n = (int)(q / g.oh);
    int ih = oh * g.stride + kh - g.pb, iw = ow * g.stride + kw - g.pb;
    if (ih < 0 || ih >= g.h || iw < 0 || iw >= g.w) continue;
    acc = fmaf(x[(((size_t)n * g.h + ih) * g.w + iw) * g.ci + ci], dy[(size_t)p * g.co + co], acc);
  }
  atomicAdd(dw + e, acc * g.alpha);
}
