// Kernels of the ResNet pitch classifier that `GANSynth.evaluate` runs real and generated spectrograms through
// (reference networks.py:293-413, ops.py:118-146 group_normalization, ops.py:308-316 max_pooling2d) and that
// models.PitchClassifier trains (models.py:253-410): NHWC fp32, forward and first-order gradients, plus the Momentum
// optimiser step.  The 3x3 convolutions of the residual blocks run on the tensor-core kernels of conv.cu; these are the
// HBM-bound rest -- plain kernels: the classifier is an evaluation network, off the benchmarked path.
#include "common.cuh"
#include "gansynth_b200.h"

namespace {

// ---- group normalisation, pass 1: per (sample, group) sum and sum of squares -------------------------------------
// x [n, hw, c]; stats [n, groups, 2] (zeroed by the host).  A thread keeps one float4 channel quad (c % 4 == 0, the quad
// never straddles a group since (c / groups) % 4 == 0 or the group is narrower than 4 -> scalar path) and walks pixels;
// the block reduces per group in shared memory, one atomic pair per (block, group).
__global__ void __launch_bounds__(256) group_stats_kernel(const float* __restrict__ x, float* __restrict__ stats, long long hw, int c,
                                                          int groups, long long pix_per_block) {
  extern __shared__ float sh[];      // [groups][2]
  const int n = blockIdx.y;
  const int cpg = c / groups;
  for (int i = threadIdx.x; i < 2 * groups; i += blockDim.x) sh[i] = 0.0f;
  __syncthreads();
  const long long p0 = blockIdx.x * pix_per_block;
  const long long p1 = p0 + pix_per_block < hw ? p0 + pix_per_block : hw;
  const float* xn = x + (size_t)n * hw * c;
  if (cpg % 4 == 0 && c / 4 <= (int)blockDim.x) {
    const int quads = c / 4;
    const int lanes = blockDim.x / quads > 0 ? blockDim.x / quads : 1;      // pixel lanes
    const int q = threadIdx.x % quads, pl = threadIdx.x / quads;
    float s = 0.0f, ss = 0.0f;
    if (pl < lanes && threadIdx.x < lanes * quads) {
      for (long long p = p0 + pl; p < p1; p += lanes) {
        const float4 v = *reinterpret_cast<const float4*>(xn + (size_t)p * c + 4 * q);
        s += (v.x + v.y) + (v.z + v.w);
        ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      }
      const int g = (4 * q) / cpg;
      atomicAdd(&sh[2 * g], s);
      atomicAdd(&sh[2 * g + 1], ss);
    }
  } else {
    for (long long i = p0 * c + threadIdx.x; i < p1 * c; i += blockDim.x) {
      const float v = xn[i];
      const int g = (int)(i % c) / cpg;
      atomicAdd(&sh[2 * g], v);
      atomicAdd(&sh[2 * g + 1], v * v);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * groups; i += blockDim.x) atomicAdd(stats + (size_t)n * 2 * groups + i, sh[i]);
}

// ---- pass 2: y = (x - mean) / sqrt(var + eps) * gamma + beta, optionally followed by relu ---------------------------
__global__ void group_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ y, long long hw, int c, int groups,
                                   float eps, int relu, size_t total) {
  const int cpg = c / groups;
  const float inv_cnt = 1.0f / ((float)hw * (float)cpg);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const size_t n = i / ((size_t)hw * c);
    const float* st = stats + (n * groups + ch / cpg) * 2;
    const float mean = st[0] * inv_cnt;
    const float var = fmaxf(st[1] * inv_cnt - mean * mean, 0.0f);        // tf.nn.moments: biased variance
    float v = (x[i] - mean) / sqrtf(var + eps) * gamma[ch] + beta[ch];
    if (relu) v = fmaxf(v, 0.0f);
    y[i] = v;
  }
}

// ---- max pooling, kernel k, stride s, TF SAME (pad_before = max(k - s, 0) / 2; padding never wins) -----------------
__global__ void max_pool_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int h, int w, int c, int k, int s,
                                int pb, int oh, int ow) {
  const size_t total = (size_t)n * oh * ow * c;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    size_t p = i / c;
    const int x0 = (int)(p % ow);
    p /= ow;
    const int y0 = (int)(p % oh);
    const int b = (int)(p / oh);
    float m = -INFINITY;
    for (int dy = 0; dy < k; ++dy) {
      const int iy = y0 * s + dy - pb;
      if (iy < 0 || iy >= h) continue;
      for (int dx = 0; dx < k; ++dx) {
        const int ix = x0 * s + dx - pb;
        if (ix < 0 || ix >= w) continue;
        m = fmaxf(m, x[(((size_t)b * h + iy) * w + ix) * c + ch]);
      }
    }
    y[i] = m;
  }
}

// ---- spatial mean: [n, hw, c] -> [n, c] (tf.reduce_mean over axes 2, 3 of NCHW); y zeroed by the host ----------------
__global__ void __launch_bounds__(256) spatial_mean_kernel(const float* __restrict__ x, float* __restrict__ y, long long hw, int c,
                                                           long long pix_per_block, float inv_hw) {
  const int n = blockIdx.y;
  const long long p0 = blockIdx.x * pix_per_block;
  const long long p1 = p0 + pix_per_block < hw ? p0 + pix_per_block : hw;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s = 0.0f;
    for (long long p = p0; p < p1; ++p) s += x[((size_t)n * hw + p) * c + ch];
    atomicAdd(y + (size_t)n * c + ch, s * inv_hw);
  }
}

// ---- group normalisation backward (training the classifier, models.py:253-304) -----------------------------------------
// With xh = (x - mean) / sigma, g = dy' * gamma (dy' = dy masked by the fused relu: y > 0), M = hw * c / groups:
//   dx = (g - mean_M(g) - xh * mean_M(g * xh)) / sigma,   dgamma[c] = sum dy' * xh,   dbeta[c] = sum dy'.
// Pass 1: per (sample, group) sums of g and g * xh into red [n, groups, 2], per-channel dgamma / dbeta (all zeroed by the
// host).  Layout and thread mapping as group_stats_kernel (scalar path only: gradients are off the hot path).
__global__ void __launch_bounds__(256) group_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                              const float* __restrict__ dy, const float* __restrict__ stats,
                                                              const float* __restrict__ gamma, float* __restrict__ red,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta, long long hw,
                                                              int c, int groups, float eps, int relu, long long pix_per_block) {
  extern __shared__ float sh[];      // [groups][2] | dgamma [c] | dbeta [c]
  float* sg = sh + 2 * groups;
  float* sb = sg + c;
  const int n = blockIdx.y;
  const int cpg = c / groups;
  const float inv_cnt = 1.0f / ((float)hw * (float)cpg);
  for (int i = threadIdx.x; i < 2 * groups + 2 * c; i += blockDim.x) sh[i] = 0.0f;
  __syncthreads();
  const long long p0 = blockIdx.x * pix_per_block;
  const long long p1 = p0 + pix_per_block < hw ? p0 + pix_per_block : hw;
  const size_t base = (size_t)n * hw * c;
  // a thread keeps one channel (c <= blockDim.x) or strides over channels; pixels strided by the lanes of that channel
  for (int ch = threadIdx.x % c; ch < c; ch += (blockDim.x >= c ? c : blockDim.x)) {
    const int lanes = blockDim.x >= c ? blockDim.x / c : 1, pl = blockDim.x >= c ? threadIdx.x / c : 0;
    if (pl >= lanes) continue;
    const int g = ch / cpg;
    const float* st = stats + ((size_t)n * groups + g) * 2;
    const float mean = st[0] * inv_cnt;
    const float rs = rsqrtf(fmaxf(st[1] * inv_cnt - mean * mean, 0.0f) + eps);
    const float gm = gamma[ch];
    float s1 = 0.0f, s2 = 0.0f, dg = 0.0f, db = 0.0f;
    for (long long p = p0 + pl; p < p1; p += lanes) {
      const size_t i = base + (size_t)p * c + ch;
      float d = dy[i];
      if (relu && !(y[i] > 0.0f)) d = 0.0f;
      const float xh = (x[i] - mean) * rs;
      s1 += d * gm;
      s2 += d * gm * xh;
      dg += d * xh;
      db += d;
    }
    atomicAdd(&sh[2 * g], s1);
    atomicAdd(&sh[2 * g + 1], s2);
    atomicAdd(&sg[ch], dg);
    atomicAdd(&sb[ch], db);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * groups; i += blockDim.x) atomicAdd(red + (size_t)n * 2 * groups + i, sh[i]);
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    atomicAdd(dgamma + i, sg[i]);
    atomicAdd(dbeta + i, sb[i]);
  }
}

__global__ void group_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                                       const float* __restrict__ stats, const float* __restrict__ gamma,
                                       const float* __restrict__ red, float* __restrict__ dx, long long hw, int c, int groups,
                                       float eps, int relu, size_t total) {
  const int cpg = c / groups;
  const float inv_cnt = 1.0f / ((float)hw * (float)cpg);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const size_t n = i / ((size_t)hw * c);
    const size_t sg = (n * groups + ch / cpg) * 2;
    const float mean = stats[sg] * inv_cnt;
    const float rs = rsqrtf(fmaxf(stats[sg + 1] * inv_cnt - mean * mean, 0.0f) + eps);
    float d = dy[i];
    if (relu && !(y[i] > 0.0f)) d = 0.0f;
    const float xh = (x[i] - mean) * rs;
    dx[i] = (d * gamma[ch] - red[sg] * inv_cnt - xh * red[sg + 1] * inv_cnt) * rs;
  }
}

// ---- max pooling backward: the gradient of a window goes to its FIRST maximum in row-major window order -------------
// (what tf.nn.max_pool's gradient does; it matters on exact ties, e.g. the constant region the zero-padded head of a clip
// produces).  Gather form: an input element scans the windows that contain it and takes a window's gradient when no
// earlier element of that window holds the same maximum.
__global__ void max_pool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                                    float* __restrict__ dx, int n, int h, int w, int c, int k, int s, int pb, int oh, int ow) {
  const size_t total = (size_t)n * h * w * c;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    size_t p = i / c;
    const int ix = (int)(p % w);
    p /= w;
    const int iy = (int)(p % h);
    const int b = (int)(p / h);
    const float v = x[i];
    float acc = 0.0f;
    // windows (oy, ox) with oy * s - pb <= iy < oy * s - pb + k
    for (int oy = max(0, (iy + pb - k + s) / s); oy < oh && oy * s - pb <= iy; ++oy)
      for (int ox = max(0, (ix + pb - k + s) / s); ox < ow && ox * s - pb <= ix; ++ox) {
        const size_t o = (((size_t)b * oh + oy) * ow + ox) * c + ch;
        if (y[o] != v) continue;
        bool first = true;                       // is there an earlier element of this window with the same value?
        for (int wy = oy * s - pb; first && wy <= iy; ++wy) {
          if (wy < 0) continue;
          const int x_end = (wy == iy) ? ix : min(w, ox * s - pb + k);
          for (int wx = max(0, ox * s - pb); wx < x_end; ++wx)
            if (x[(((size_t)b * h + wy) * w + wx) * c + ch] == v) { first = false; break; }
        }
        if (first) acc += dy[o];
      }
    dx[i] = acc;
  }
}

// ---- spatial mean backward: dx[n, p, c] = dy[n, c] / hw ---------------------------------------------------------------
__global__ void spatial_mean_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long long hw, int c, float inv_hw,
                                        size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / ((size_t)hw * c);
    dx[i] = dy[n * c + i % c] * inv_hw;
  }
}

// ---- tf.train.MomentumOptimizer (models.py:283-295), weight decay of models.py:267-271 folded in as wd * p ------------
// accum = momentum * accum + g;  p -= lr * (nesterov ? g + momentum * accum : accum)      (g includes wd * p where wd != 0)
__global__ void momentum_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ accum,
                                const float* __restrict__ wd, size_t n, float lr, float momentum, int nesterov, float gscale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale + (wd ? wd[i] * p[i] : 0.0f);
    const float a = momentum * accum[i] + gi;
    accum[i] = a;
    p[i] -= lr * (nesterov ? gi + momentum * a : a);
  }
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int gs_group_norm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* stats, int n,
                                 long long hw, int c, int groups, float eps, int relu, void* stream) {
  GS_CHECK_ARG(n > 0 && hw > 0 && c > 0 && groups > 0 && c % groups == 0, "group_norm: bad shape (c %d, groups %d)", c, groups);
  GS_CHECK_ARG(groups <= 1024, "group_norm: at most 1024 groups");
  GS_CUDA(cudaMemsetAsync(stats, 0, (size_t)n * groups * 2 * sizeof(float), ST));
  long long blocks_x = (2LL * gs_num_sms() + n - 1) / n;
  long long ppb = (hw + blocks_x - 1) / blocks_x;
  if (ppb < 16) ppb = 16;
  blocks_x = (hw + ppb - 1) / ppb;
  group_stats_kernel<<<dim3((unsigned)blocks_x, (unsigned)n), 256, (size_t)2 * groups * sizeof(float), ST>>>(x, stats, hw, c, groups, ppb);
  GS_CHECK_LAUNCH("group_stats");
  const size_t total = (size_t)n * hw * c;
  size_t b = (total + 1023) / 1024;
  const size_t cap = (size_t)gs_num_sms() * 16;
  group_apply_kernel<<<(unsigned)(b < cap ? b : cap), 256, 0, ST>>>(x, stats, gamma, beta, y, hw, c, groups, eps, relu, total);
  GS_CHECK_LAUNCH("group_apply");
  return GS_OK;
}

extern "C" int gs_max_pool2d(const float* x, float* y, int n, int h, int w, int c, int ksize, int stride, void* stream) {
  GS_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && ksize >= 1 && stride >= 1, "max_pool2d: bad shape");
  const int oh = (h + stride - 1) / stride, ow = (w + stride - 1) / stride;       // SAME
  // TF SAME: total padding max((o - 1) * s + k - size, 0), the smaller half in front
  const int pad_h = (oh - 1) * stride + ksize - h, pad_w = (ow - 1) * stride + ksize - w;
  GS_CHECK_ARG((pad_h > 0 ? pad_h : 0) / 2 == (pad_w > 0 ? pad_w : 0) / 2, "max_pool2d: unequal SAME padding in h and w is not supported");
  const int pb = (pad_h > 0 ? pad_h : 0) / 2;
  const size_t total = (size_t)n * oh * ow * c;
  size_t b = (total + 255) / 256;
  const size_t cap = (size_t)gs_num_sms() * 16;
  max_pool_kernel<<<(unsigned)(b < cap ? b : cap), 256, 0, ST>>>(x, y, n, h, w, c, ksize, stride, pb, oh, ow);
  GS_CHECK_LAUNCH("max_pool2d");
  return GS_OK;
}

extern "C" int gs_spatial_mean(const float* x, float* y, int n, long long hw, int c, void* stream) {
  GS_CHECK_ARG(n > 0 && hw > 0 && c > 0, "spatial_mean: bad shape");
  GS_CUDA(cudaMemsetAsync(y, 0, (size_t)n * c * sizeof(float), ST));
  long long blocks_x = (2LL * gs_num_sms() + n - 1) / n;
  long long ppb = (hw + blocks_x - 1) / blocks_x;
  if (ppb < 8) ppb = 8;
  blocks_x = (hw + ppb - 1) / ppb;
  spatial_mean_kernel<<<dim3((unsigned)blocks_x, (unsigned)n), 256, 0, ST>>>(x, y, hw, c, ppb, 1.0f / (float)hw);
  GS_CHECK_LAUNCH("spatial_mean");
  return GS_OK;
}


static unsigned cls_grid(size_t total, int per_block = 1024) {
  size_t b = (total + per_block - 1) / per_block;
  const size_t cap = (size_t)gs_num_sms() * 16;
  return (unsigned)(b < cap ? (b ? b : 1) : cap);
}

// x, y (the forward output, read only for the relu mask), dy, stats (of the forward) -> dx, dgamma [c], dbeta [c];
// red: caller scratch of n * groups * 2 floats
extern "C" int gs_group_norm_bwd(const float* x, const float* y, const float* dy, const float* stats, const float* gamma,
                                 float* dx, float* dgamma, float* dbeta, float* red, int n, long long hw, int c, int groups,
                                 float eps, int relu, void* stream) {
  GS_CHECK_ARG(n > 0 && hw > 0 && c > 0 && groups > 0 && c % groups == 0 && groups <= 1024 && c <= 4096,
               "group_norm_bwd: bad shape (c %d, groups %d)", c, groups);
  GS_CUDA(cudaMemsetAsync(red, 0, (size_t)n * groups * 2 * sizeof(float), ST));
  GS_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)c * sizeof(float), ST));
  GS_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)c * sizeof(float), ST));
  long long blocks_x = (2LL * gs_num_sms() + n - 1) / n;
  long long ppb = (hw + blocks_x - 1) / blocks_x;
  if (ppb < 16) ppb = 16;
  blocks_x = (hw + ppb - 1) / ppb;
  const size_t smem = (size_t)(2 * groups + 2 * c) * sizeof(float);
  group_bwd_stats_kernel<<<dim3((unsigned)blocks_x, (unsigned)n), 256, smem, ST>>>(x, y, dy, stats, gamma, red, dgamma, dbeta, hw, c, groups,
                                                                                    eps, relu, ppb);
  GS_CHECK_LAUNCH("group_bwd_stats");
  const size_t total = (size_t)n * hw * c;
  group_bwd_apply_kernel<<<cls_grid(total), 256, 0, ST>>>(x, y, dy, stats, gamma, red, dx, hw, c, groups, eps, relu, total);
  GS_CHECK_LAUNCH("group_bwd_apply");
  return GS_OK;
}

extern "C" int gs_max_pool2d_bwd(const float* x, const float* y, const float* dy, float* dx, int n, int h, int w, int c,
                                 int ksize, int stride, void* stream) {
  GS_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && ksize >= 1 && stride >= 1, "max_pool2d_bwd: bad shape");
  const int oh = (h + stride - 1) / stride, ow = (w + stride - 1) / stride;
  const int pad_h = (oh - 1) * stride + ksize - h;
  const int pb = (pad_h > 0 ? pad_h : 0) / 2;
  max_pool_bwd_kernel<<<cls_grid((size_t)n * h * w * c, 256), 256, 0, ST>>>(x, y, dy, dx, n, h, w, c, ksize, stride, pb, oh, ow);
  GS_CHECK_LAUNCH("max_pool2d_bwd");
  return GS_OK;
}

extern "C" int gs_spatial_mean_bwd(const float* dy, float* dx, int n, long long hw, int c, void* stream) {
  GS_CHECK_ARG(n > 0 && hw > 0 && c > 0, "spatial_mean_bwd: bad shape");
  const size_t total = (size_t)n * hw * c;
  spatial_mean_bwd_kernel<<<cls_grid(total), 256, 0, ST>>>(dy, dx, hw, c, 1.0f / (float)hw, total);
  GS_CHECK_LAUNCH("spatial_mean_bwd");
  return GS_OK;
}

// wd: per-element weight-decay coefficient [n] (0 for the normalisation variables) or NULL
extern "C" int gs_momentum_step(float* p, const float* g, float* accum, const float* wd, long long n, float lr, float momentum,
                                int nesterov, float grad_scale, void* stream) {
  GS_CHECK_ARG(n >= 0, "momentum_step: negative size");
  if (n == 0) return GS_OK;
  momentum_kernel<<<cls_grid((size_t)n), 256, 0, ST>>>(p, g, accum, wd, (size_t)n, lr, momentum, nesterov, grad_scale);
  GS_CHECK_LAUNCH("momentum_step");
  return GS_OK;
}
