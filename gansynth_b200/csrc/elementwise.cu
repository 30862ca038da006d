// HBM-bound per-pixel / elementwise kernels of the GANSynth step (NHWC fp32) and their first- and
// second-order derivative forms.  Reference call sites: pixel_normalization ops.py:330-333,
// batch_stddev ops.py:336-348, upscale2d/downscale2d ops.py:283-305, lerp networks.py:10-11,
// tf.nn.leaky_relu / tf.nn.tanh in networks.py, tf.train.AdamOptimizer models.py:67-76.
#include "common.cuh"
#include "gansynth_b200.h"

namespace {

constexpr int EW_BLOCK = 256;

inline int ew_grid(size_t n, int per_thread = 4) {
  size_t b = (n + (size_t)EW_BLOCK * per_thread - 1) / ((size_t)EW_BLOCK * per_thread);
  size_t cap = (size_t)gs_num_sms() * 16;
  if (b < 1) b = 1;
  return (int)(b < cap ? b : cap);
}

#define GS_GRID_STRIDE(i, n) \
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (n); i += (size_t)gridDim.x * blockDim.x)

// ---------------------------------------------------------------- simple elementwise
__global__ void mask_mul_kernel(const float* __restrict__ v, const float* __restrict__ y, float* __restrict__ o, size_t n) {
  GS_GRID_STRIDE(i, n) o[i] = v[i] * gs_lrelu_slope(y[i]);
}
__global__ void mask_mul4_kernel(const float4* __restrict__ v, const float4* __restrict__ y, float4* __restrict__ o, size_t n4) {
  GS_GRID_STRIDE(i, n4) {
    float4 a = v[i], b = y[i], r;
    r.x = a.x * gs_lrelu_slope(b.x); r.y = a.y * gs_lrelu_slope(b.y);
    r.z = a.z * gs_lrelu_slope(b.z); r.w = a.w * gs_lrelu_slope(b.w);
    o[i] = r;
  }
}
// o = v * lrelu'(y) and colsum[c] += sum_rows o  (bias gradient of the layer in the same pass).  The grid stride is
// a multiple of c / 4, so every thread stays on one group of four channels.
__global__ void mask_mul_colsum_kernel(const float4* __restrict__ v, const float4* __restrict__ y, float4* __restrict__ o,
                                       float* __restrict__ colsum, size_t n4, int c4) {
  __shared__ float cs[256];
  for (int i = threadIdx.x; i < 4 * c4; i += blockDim.x) cs[i] = 0.0f;
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  GS_GRID_STRIDE(i, n4) {
    float4 a = v[i], b = y[i], r;
    r.x = a.x * gs_lrelu_slope(b.x); r.y = a.y * gs_lrelu_slope(b.y);
    r.z = a.z * gs_lrelu_slope(b.z); r.w = a.w * gs_lrelu_slope(b.w);
    o[i] = r;
    acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
  }
  const int g = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) % (size_t)c4);
  atomicAdd(&cs[4 * g + 0], acc.x); atomicAdd(&cs[4 * g + 1], acc.y);
  atomicAdd(&cs[4 * g + 2], acc.z); atomicAdd(&cs[4 * g + 3], acc.w);
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * c4; i += blockDim.x) atomicAdd(colsum + i, cs[i]);
}
__global__ void lrelu_kernel(const float* __restrict__ x, float* __restrict__ o, size_t n) {
  GS_GRID_STRIDE(i, n) o[i] = gs_lrelu(x[i]);
}
__global__ void tanh_kernel(const float* __restrict__ x, float* __restrict__ o, size_t n) {
  GS_GRID_STRIDE(i, n) o[i] = tanhf(x[i]);
}
// dz = dy * (1 - y^2)
__global__ void tanh_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ o, size_t n) {
  GS_GRID_STRIDE(i, n) { float t = y[i]; o[i] = dy[i] * (1.0f - t * t); }
}
// gy = -2 * y * dy * u
__global__ void tanh_bwd2_kernel(const float* __restrict__ y, const float* __restrict__ dy, const float* __restrict__ u,
                                 float* __restrict__ o, size_t n) {
  GS_GRID_STRIDE(i, n) o[i] = -2.0f * y[i] * dy[i] * u[i];
}
// o = alpha * a + beta * b   (b may be null)
__global__ void axpby_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, float alpha,
                             float beta, size_t n) {
  GS_GRID_STRIDE(i, n) o[i] = b ? fmaf(alpha, a[i], beta * b[i]) : alpha * a[i];
}
// o = coef[ia] * a + coef[ib] * b with the coefficients in DEVICE memory (b may be null: o = coef[ia] * a).  The
// progressive-growing blend weight changes every step; read from memory it can live outside a replayed CUDA graph.
__global__ void axpby_dev_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o,
                                 const float* __restrict__ coef, int ia, int ib, size_t n) {
  const float alpha = __ldg(coef + ia), beta = __ldg(coef + ib);
  GS_GRID_STRIDE(i, n) o[i] = b ? fmaf(alpha, a[i], beta * b[i]) : alpha * a[i];
}
__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, float alpha, size_t n) {
  GS_GRID_STRIDE(i, n) o[i] = alpha * a[i] * b[i];
}
// o[r,c] = act(x[r,c] + bias[c])
__global__ void bias_act_kernel(const float* __restrict__ x, const float* __restrict__ bias, float* __restrict__ o,
                                size_t n, int c, int act) {
  GS_GRID_STRIDE(i, n) {
    float v = x[i] + (bias ? bias[i % c] : 0.0f);
    o[i] = act == 1 ? gs_lrelu(v) : v;
  }
}
// o[r,c] = s[c]
__global__ void row_broadcast_kernel(const float* __restrict__ s, float* __restrict__ o, size_t n, int c) {
  GS_GRID_STRIDE(i, n) o[i] = s[i % c];
}

// ---------------------------------------------------------------- column sum  out[c] = sum_r v[r,c]
__global__ void col_sum_kernel(const float* __restrict__ v, float* __restrict__ out, long long rows, int c,
                               long long rows_per_block) {
  extern __shared__ float red[];
  long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  int lanes = blockDim.x / c;
  if (lanes < 1) lanes = 1;
  for (int cb = 0; cb < c; cb += blockDim.x) {
    int tid = threadIdx.x;
    int lane = (lanes > 1) ? tid / c : 0;
    int col = (lanes > 1) ? tid % c : cb + tid;
    float acc = 0.0f;
    if (lane < lanes && col < c)
      for (long long r = r0 + lane; r < r1; r += lanes) acc += v[(size_t)r * c + col];
    if (lanes > 1) {
      red[tid] = acc;
      __syncthreads();
      if (tid < c) {
        float s = 0.0f;
        for (int l = 0; l < lanes; ++l) s += red[l * c + tid];
        atomicAdd(out + tid, s);
      }
      __syncthreads();
    } else if (col < c) {
      atomicAdd(out + col, acc);
    }
  }
}

// ---------------------------------------------------------------- pixel norm (per-row L2 normalisation over C)
// One warp per row; lanes stride over float4 groups.  r = 1/sqrt(mean(a^2)+eps), y = a*r.
template <int MODE>  // 0 fwd, 1 bwd, 2 bwd2
__global__ void pixel_norm_kernel(const float* __restrict__ a, const float* __restrict__ rin, const float* __restrict__ dy,
                                  const float* __restrict__ u, float* __restrict__ out, float* __restrict__ rout,
                                  long long rows, int c, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float inv_c = 1.0f / (float)c;
  const bool vec = (c % 4) == 0;
  float4 cacc0 = make_float4(0.f, 0.f, 0.f, 0.f), cacc1 = cacc0;   // MODE 3: bias-gradient partial sums of this lane
  (void)cacc0; (void)cacc1;
  for (long long row = warp; row < rows; row += nwarps) {
    const float* ap = a + (size_t)row * c;
    float* op = out + (size_t)row * c;
    if (MODE == 0) {
      float ss = 0.0f;
      if (vec) for (int i = lane * 4; i < c; i += 128) { float4 t = *reinterpret_cast<const float4*>(ap + i); ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w; }
      else for (int i = lane; i < c; i += 32) ss += ap[i] * ap[i];
      ss = gs_warp_sum(ss);
      float r = 1.0f / sqrtf(ss * inv_c + eps);
      if (vec) for (int i = lane * 4; i < c; i += 128) { float4 t = *reinterpret_cast<const float4*>(ap + i); t.x *= r; t.y *= r; t.z *= r; t.w *= r; *reinterpret_cast<float4*>(op + i) = t; }
      else for (int i = lane; i < c; i += 32) op[i] = ap[i] * r;
      if (lane == 0) rout[row] = r;
    } else if (MODE == 1) {
      // da = r*dy - (r^3/C) (a.dy) a
      const float* dp = dy + (size_t)row * c;
      float dot = 0.0f;
      if (vec) for (int i = lane * 4; i < c; i += 128) { float4 t = *reinterpret_cast<const float4*>(ap + i); float4 d = *reinterpret_cast<const float4*>(dp + i); dot += t.x * d.x + t.y * d.y + t.z * d.z + t.w * d.w; }
      else for (int i = lane; i < c; i += 32) dot += ap[i] * dp[i];
      dot = gs_warp_sum(dot);
      float r = rin[row];
      float k = r * r * r * inv_c * dot;
      if (vec) for (int i = lane * 4; i < c; i += 128) {
        float4 t = *reinterpret_cast<const float4*>(ap + i); float4 d = *reinterpret_cast<const float4*>(dp + i); float4 o;
        o.x = r * d.x - k * t.x; o.y = r * d.y - k * t.y; o.z = r * d.z - k * t.z; o.w = r * d.w - k * t.w;
        *reinterpret_cast<float4*>(op + i) = o; }
      else for (int i = lane; i < c; i += 32) op[i] = r * dp[i] - k * ap[i];
    } else if (MODE == 3) {
      // dz = lrelu'(a) * [ r*dy - (r^3/C) (a.dy) a ]: pixel-norm backward and the leaky-relu mask of the layer
      // that produced a, in one pass; optionally the bias gradient colsum[c] += sum_rows dz (c % 4 == 0, c <= 256)
      const float* dp = dy + (size_t)row * c;
      float dot = 0.0f;
      for (int i = lane * 4; i < c; i += 128) { float4 t = *reinterpret_cast<const float4*>(ap + i); float4 d = *reinterpret_cast<const float4*>(dp + i); dot += t.x * d.x + t.y * d.y + t.z * d.z + t.w * d.w; }
      dot = gs_warp_sum(dot);
      float r = rin[row];
      float k = r * r * r * inv_c * dot;
      int slot = 0;
      for (int i = lane * 4; i < c; i += 128, ++slot) {
        float4 t = *reinterpret_cast<const float4*>(ap + i); float4 d = *reinterpret_cast<const float4*>(dp + i); float4 o;
        o.x = (r * d.x - k * t.x) * gs_lrelu_slope(t.x); o.y = (r * d.y - k * t.y) * gs_lrelu_slope(t.y);
        o.z = (r * d.z - k * t.z) * gs_lrelu_slope(t.z); o.w = (r * d.w - k * t.w) * gs_lrelu_slope(t.w);
        *reinterpret_cast<float4*>(op + i) = o;
        if (slot == 0) { cacc0.x += o.x; cacc0.y += o.y; cacc0.z += o.z; cacc0.w += o.w; }
        else { cacc1.x += o.x; cacc1.y += o.y; cacc1.z += o.z; cacc1.w += o.w; }
      }
    } else {
      // ga = -(r^3/C) [ (u.dy) a + (a.dy) u + (u.a) dy ] + 3 (r^5/C^2) (u.a)(a.dy) a
      const float* dp = dy + (size_t)row * c;
      const float* up = u + (size_t)row * c;
      float ud = 0.0f, ad = 0.0f, ua = 0.0f;
      for (int i = lane; i < c; i += 32) { float t = ap[i], d = dp[i], w = up[i]; ud += w * d; ad += t * d; ua += w * t; }
      ud = gs_warp_sum(ud); ad = gs_warp_sum(ad); ua = gs_warp_sum(ua);
      float r = rin[row];
      float r3c = r * r * r * inv_c;
      float ka = -r3c * ud + 3.0f * r3c * r * r * inv_c * ua * ad;
      float ku = -r3c * ad, kd = -r3c * ua;
      for (int i = lane; i < c; i += 32) op[i] = ka * ap[i] + ku * up[i] + kd * dp[i];
    }
  }
  if (MODE == 3 && rout != nullptr) {
    // rout doubles as the bias-gradient output [c] in this mode
    __shared__ float cs[256];
    for (int i = threadIdx.x; i < c; i += blockDim.x) cs[i] = 0.0f;
    __syncthreads();
    if (lane * 4 < c) { atomicAdd(&cs[lane * 4], cacc0.x); atomicAdd(&cs[lane * 4 + 1], cacc0.y); atomicAdd(&cs[lane * 4 + 2], cacc0.z); atomicAdd(&cs[lane * 4 + 3], cacc0.w); }
    if (lane * 4 + 128 < c) { atomicAdd(&cs[lane * 4 + 128], cacc1.x); atomicAdd(&cs[lane * 4 + 129], cacc1.y); atomicAdd(&cs[lane * 4 + 130], cacc1.z); atomicAdd(&cs[lane * 4 + 131], cacc1.w); }
    __syncthreads();
    for (int i = threadIdx.x; i < c; i += blockDim.x) atomicAdd(rout + i, cs[i]);
  }
}

// ---------------------------------------------------------------- resampling (NHWC)
// out[n, y, x, c] = in[n, y/fh, x/fw, c] * scale        (nearest upscale; scale=1)
__global__ void upscale_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int h, int w, int c, int fh,
                               int fw, float scale) {
  size_t total = (size_t)n * h * fh * w * fw * c;
  GS_GRID_STRIDE(i, total) {
    int ch = (int)(i % c);
    size_t p = i / c;
    int x = (int)(p % (w * fw));
    p /= (w * fw);
    int y = (int)(p % (h * fh));
    int b = (int)(p / (h * fh));
    out[i] = scale * in[(((size_t)b * h + y / fh) * w + x / fw) * c + ch];
  }
}
// out[n, y, x, c] = scale * sum_{dy<fh, dx<fw} in[n, y*fh+dy, x*fw+dx, c]   (in is [n, h*fh, w*fw, c])
__global__ void pool_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int h, int w, int c, int fh,
                            int fw, float scale) {
  size_t total = (size_t)n * h * w * c;
  GS_GRID_STRIDE(i, total) {
    int ch = (int)(i % c);
    size_t p = i / c;
    int x = (int)(p % w);
    p /= w;
    int y = (int)(p % h);
    int b = (int)(p / h);
    float s = 0.0f;
    for (int dy = 0; dy < fh; ++dy)
      for (int dx = 0; dx < fw; ++dx)
        s += in[(((size_t)b * h * fh + y * fh + dy) * (w * fw) + x * fw + dx) * c + ch];
    out[i] = s * scale;
  }
}
// generic 3-axis permutation [n, a, b] -> [n, b, a]
__global__ void transpose_inner_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int a, int b) {
  size_t total = (size_t)n * a * b;
  GS_GRID_STRIDE(i, total) {
    int ia = (int)(i % a);
    size_t p = i / a;
    int ib = (int)(p % b);
    int in_ = (int)(p / b);
    out[i] = in[((size_t)in_ * a + ia) * b + ib];
  }
}

// ---------------------------------------------------------------- per-sample row ops on [B, E]
__global__ void row_dot_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, size_t e) {
  __shared__ float red[32];
  const float* ap = a + (size_t)blockIdx.y * e;
  const float* bp = b + (size_t)blockIdx.y * e;
  float s = 0.0f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < e; i += (size_t)gridDim.x * blockDim.x) s += ap[i] * bp[i];
  s = gs_warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
    t = gs_warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(out + blockIdx.y, t);
  }
}
__global__ void row_scale_kernel(const float* __restrict__ a, const float* __restrict__ s, float* __restrict__ out, size_t e,
                                 float alpha) {
  const float k = alpha * s[blockIdx.y];
  const float* ap = a + (size_t)blockIdx.y * e;
  float* op = out + (size_t)blockIdx.y * e;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < e; i += (size_t)gridDim.x * blockDim.x) op[i] = k * ap[i];
}

// ---------------------------------------------------------------- minibatch stddev (ops.py:336-348), NHWC [B, E]
// groups g = 0..G-1, members m = 0..M-1 (M = B/G), sample index = g*M + m, E = H*W*C elements.
// MODE 0: stat[m] += (1/E) sum_e sqrt(var_g + eps)
// MODE 1: dx[g,m,e] = df[m] * (x - mu) / (E G s)
// MODE 2: gx[h,m,e] = c_m * ( (u_h - ubar)/s - (sum_g u_g (x_g - mu)) (x_h - mu) / (G s^3) ),  c_m = df[m]/(E G);
//         q[m] += (1/(E G)) sum_e sum_g u_g (x_g - mu) / s
template <int MODE>
__global__ void stddev_kernel(const float* __restrict__ x, const float* __restrict__ df, const float* __restrict__ u,
                              float* __restrict__ out, float* __restrict__ stat, int G, int M, size_t E, float eps) {
  __shared__ float red[32];
  const int m = blockIdx.y;
  float local = 0.0f;
  const float invG = 1.0f / (float)G, invE = 1.0f / (float)E;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < E; e += (size_t)gridDim.x * blockDim.x) {
    float mu = 0.0f;
    for (int g = 0; g < G; ++g) mu += x[((size_t)(g * M + m)) * E + e];
    mu *= invG;
    float var = 0.0f;
    for (int g = 0; g < G; ++g) { float d = x[((size_t)(g * M + m)) * E + e] - mu; var += d * d; }
    var *= invG;
    float s = sqrtf(var + eps);
    if (MODE == 0) {
      local += s;
    } else if (MODE == 1) {
      float k = df[m] * invE * invG / s;
      for (int g = 0; g < G; ++g) { size_t o = ((size_t)(g * M + m)) * E + e; out[o] = k * (x[o] - mu); }
    } else {
      float ubar = 0.0f, ux = 0.0f;
      for (int g = 0; g < G; ++g) { size_t o = ((size_t)(g * M + m)) * E + e; ubar += u[o]; ux += u[o] * (x[o] - mu); }
      ubar *= invG;
      float cm = df[m] * invE * invG;
      float k2 = ux * invG / (s * s * s);
      for (int g = 0; g < G; ++g) { size_t o = ((size_t)(g * M + m)) * E + e; out[o] = cm * ((u[o] - ubar) / s - k2 * (x[o] - mu)); }
      local += ux / s;
    }
  }
  if (MODE != 1) {
    local = gs_warp_sum(local);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
      t = gs_warp_sum(t);
      if (threadIdx.x == 0) atomicAdd(stat + m, t * invE * (MODE == 2 ? invG : 1.0f));
    }
  }
}

// ---------------------------------------------------------------- embedding (ops.py:204-218)
__global__ void embedding_fwd_kernel(const float* __restrict__ table, const long long* __restrict__ idx, float* __restrict__ out,
                                     int b, int units, float alpha) {
  size_t total = (size_t)b * units;
  GS_GRID_STRIDE(i, total) out[i] = alpha * table[(size_t)idx[i / units] * units + (i % units)];
}
__global__ void embedding_bwd_kernel(const float* __restrict__ dy, const long long* __restrict__ idx, float* __restrict__ dtable,
                                     int b, int units, float alpha) {
  size_t total = (size_t)b * units;
  GS_GRID_STRIDE(i, total) atomicAdd(dtable + (size_t)idx[i / units] * units + (i % units), alpha * dy[i]);
}

// ---------------------------------------------------------------- TF Adam
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            size_t n, float lr_t, float b1, float b2, float eps, float gscale) {
  GS_GRID_STRIDE(i, n) {
    float gi = g[i] * gscale;
    float mi = b1 * m[i] + (1.0f - b1) * gi;
    float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

// ---------------------------------------------------------------- gradient all-reduce FUSED with TF Adam over NVLink
// Data-parallel update of models.py:81-89 as ONE kernel per network instead of ncclAllReduce + Adam: rank r owns the
// slice [lo, hi) of the flat buffers.  It (1) reads the SUM of all ranks' gradients of its slice -- MODE 0: one
// multimem.ld_reduce per 16 bytes on the NVSwitch multicast address (the switch adds the eight copies in flight);
// MODE 1: plain peer loads through the NVLink-mapped pointers -- (2) applies TF-Adam to its slice (the Adam slots m, v
// are SHARDED: only the owner's slice is ever valid), (3) writes the new parameters into EVERY rank's buffer
// (multimem.st broadcast / peer stores).  Per rank and sub-step: n/W * 4 bytes in, n * 4 * (W-1)/W bytes out of the
// switch, against 2 * n * 4 * (W-1)/W each way for a ring all-reduce, and Adam touches n/W elements instead of n.
// The caller brackets the launch with cross-rank barriers (all gradients written / all parameters visible).
__device__ __forceinline__ float4 mm_ld_reduce_f32x4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void mm_st_f32x4(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

template <int MODE>
__global__ void adam_reduce_kernel(const float* __restrict__ p_local, float* __restrict__ m, float* __restrict__ v,
                                   const float* g_mc, float* p_mc, const float* const* __restrict__ g_peers,
                                   float* const* __restrict__ p_peers, int world, long long lo4, long long hi4, float lr_t,
                                   float b1, float b2, float eps, float gscale) {
  for (long long q = lo4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; q < hi4; q += (long long)gridDim.x * blockDim.x) {
    const long long i = 4 * q;
    float4 g;
    if (MODE == 0) {
      g = mm_ld_reduce_f32x4(g_mc + i);
    } else {
      g = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < world; ++r) {
        const float4 t = *reinterpret_cast<const float4*>(g_peers[r] + i);
        g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
      }
    }
    const float4 pm = *reinterpret_cast<const float4*>(m + i), pv = *reinterpret_cast<const float4*>(v + i);
    const float4 pp = *reinterpret_cast<const float4*>(p_local + i);
    float4 nm, nv, np;
#define GS_ADAM1(c)                                   \
    {                                                 \
      const float gi = g.c * gscale;                  \
      nm.c = b1 * pm.c + (1.0f - b1) * gi;            \
      nv.c = b2 * pv.c + (1.0f - b2) * gi * gi;       \
      np.c = pp.c - lr_t * nm.c / (sqrtf(nv.c) + eps); \
    }
    GS_ADAM1(x) GS_ADAM1(y) GS_ADAM1(z) GS_ADAM1(w)
#undef GS_ADAM1
    *reinterpret_cast<float4*>(m + i) = nm;
    *reinterpret_cast<float4*>(v + i) = nv;
    if (MODE == 0) {
      mm_st_f32x4(p_mc + i, np);
    } else {
      for (int r = 0; r < world; ++r) *reinterpret_cast<float4*>(p_peers[r] + i) = np;
    }
  }
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int gs_lrelu_mask_mul(const float* v, const float* y, float* out, long long n, void* stream) {
  GS_CHECK_ARG(n >= 0, "mask_mul: negative size");
  if (n == 0) return GS_OK;
  bool al = ((((uintptr_t)v | (uintptr_t)y | (uintptr_t)out) & 15) == 0) && (n % 4 == 0);
  if (al) mask_mul4_kernel<<<ew_grid(n / 4), EW_BLOCK, 0, ST>>>((const float4*)v, (const float4*)y, (float4*)out, n / 4);
  else mask_mul_kernel<<<ew_grid(n), EW_BLOCK, 0, ST>>>(v, y, out, n);
  GS_CHECK_LAUNCH("mask_mul");
  return GS_OK;
}
extern "C" int gs_lrelu(const float* x, float* out, long long n, void* stream) {
  if (n <= 0) return GS_OK;
  lrelu_kernel<<<ew_grid(n), EW_BLOCK, 0, ST>>>(x, out, n);
  GS_CHECK_LAUNCH("lrelu");
  return GS_OK;
}
extern "C" int gs_tanh_fwd(const float* x, float* out, long long n, void* stream) {
  if (n <= 0) return GS_OK;
  tanh_kernel<<<ew_grid(n), EW_BLOCK, 0, ST>>>(x, out, n);
  GS_CHECK_LAUNCH("tanh");
  return GS_OK;
}
extern "C" int gs_tanh_bwd(const float* y, const float* dy, float* out, long long n, void* stream) {
  if (n <= 0) return GS_OK;
  tanh_bwd_kernel<<<ew_grid(n), EW_BLOCK, 0, ST>>>(y, dy, out, n);
  GS_CHECK_LAUNCH("tanh_bwd");
  return GS_OK;
}
extern "C" int gs_tanh_bwd2(const float* y, const float* dy, const float* u, float* out, long long n, void* stream) {
  if (n <= 0) return GS_OK;
  tanh_bwd2_kernel<<<ew_grid(n), EW_BLOCK, 0, ST>>>(y, dy, u, out, n);
  GS_CHECK_LAUNCH("tanh_bwd2");
  return GS_OK;
}
extern "C" int gs_axpby(const float* a, const float* b, float* out, float alpha, float beta, long long n, void* stream) {
  if (n <= 0) return GS_OK;
  axpby_kernel<<<ew_grid(n), EW_BLOCK, 0, ST>>>(a, b, out, alpha, beta, n);
  GS_CHECK_LAUNCH("axpby");
  return GS_OK;
}
extern "C" int gs_axpby_dev(const float* a, const float* b, float* out, const float* coef, int ia, int ib, long long n,
                            void* stream) {
  GS_CHECK_ARG(coef != nullptr && ia >= 0 && ib >= 0, "axpby_dev: bad coefficients");
  if (n <= 0) return GS_OK;
  axpby_dev_kernel<<<ew_grid(n), EW_BLOCK, 0, ST>>>(a, b, out, coef, ia, ib, n);
  GS_CHECK_LAUNCH("axpby_dev");
  return GS_OK;
}
extern "C" int gs_mul(const float* a, const float* b, float* out, float alpha, long long n, void* stream) {
  if (n <= 0) return GS_OK;
  mul_kernel<<<ew_grid(n), EW_BLOCK, 0, ST>>>(a, b, out, alpha, n);
  GS_CHECK_LAUNCH("mul");
  return GS_OK;
}
extern "C" int gs_bias_act(const float* x, const float* bias, float* out, long long rows, int c, int act, void* stream) {
  GS_CHECK_ARG(rows >= 0 && c > 0, "bias_act: bad shape");
  size_t n = (size_t)rows * c;
  if (n == 0) return GS_OK;
  bias_act_kernel<<<ew_grid(n), EW_BLOCK, 0, ST>>>(x, bias, out, n, c, act);
  GS_CHECK_LAUNCH("bias_act");
  return GS_OK;
}
extern "C" int gs_row_broadcast(const float* s, float* out, long long rows, int c, void* stream) {
  size_t n = (size_t)rows * c;
  if (n == 0) return GS_OK;
  row_broadcast_kernel<<<ew_grid(n), EW_BLOCK, 0, ST>>>(s, out, n, c);
  GS_CHECK_LAUNCH("row_broadcast");
  return GS_OK;
}
int gs_col_sum_acc(const float* v, float* out, long long rows, int c, void* stream);
extern "C" int gs_col_sum(const float* v, float* out, long long rows, int c, void* stream) {
  GS_CHECK_ARG(rows >= 0 && c > 0, "col_sum: bad shape");
  GS_CUDA(cudaMemsetAsync(out, 0, (size_t)c * sizeof(float), ST));
  return gs_col_sum_acc(v, out, rows, c, stream);
}
// out[c] += column sums (library-internal: the accumulate form of gs_conv2d_wgrad_ex's fallback)
int gs_col_sum_acc(const float* v, float* out, long long rows, int c, void* stream) {
  GS_CHECK_ARG(rows >= 0 && c > 0, "col_sum: bad shape");
  if (rows == 0) return GS_OK;
  int lanes = EW_BLOCK / c;
  if (lanes < 1) lanes = 1;
  long long target_blocks = (long long)gs_num_sms() * 8;
  long long rpb = (rows + target_blocks - 1) / target_blocks;
  long long min_rpb = lanes * 4;
  if (rpb < min_rpb) rpb = min_rpb;
  int blocks = gs_cdiv(rows, rpb);
  col_sum_kernel<<<blocks, EW_BLOCK, EW_BLOCK * sizeof(float), ST>>>(v, out, rows, c, rpb);
  GS_CHECK_LAUNCH("col_sum");
  return GS_OK;
}

// Vectorised pixel-norm family: LPP = min(32, c/4) lanes per pixel, 32 / LPP pixels per warp (the 32-channel
// top-resolution activations use all 32 lanes on 4 pixels instead of 8 lanes on one), float4 accesses, reductions
// by xor-shuffles inside the LPP-lane group.  Same MODE meaning as pixel_norm_kernel; c % 4 == 0, c/4 a power of two
// <= 32 or a multiple of 32 up to 128 (c <= 512).
template <int MODE, int SLOTS, int R>
__global__ void pixel_norm_vec_kernel(const float* __restrict__ a, const float* __restrict__ rin, const float* __restrict__ dy,
                                      const float* __restrict__ u, float* __restrict__ out, float* __restrict__ rout,
                                      long long rows, int c, float eps, int lpp, int flags) {
  // SLOTS = float4 per lane per pixel ((c/4) / lpp); R = pixels per lane group in flight per iteration (memory-level
  // parallelism: every load of the R pixels is issued before the first reduction).
  // flags (second-order forms of the fused pixel-norm/leaky-relu backward): 1 = multiply the incoming vector
  // (dy in MODE 1, u in MODE 2) by lrelu'(a) first; 2 = multiply the result by lrelu'(a); 4 = `a` is given as y = a * r;
  // 8 (MODE 2) = also write the MODE 1 result for the same incoming u to `rout` (shaped like `out`)
  const int lane = threadIdx.x & 31;
  const int li = lane % lpp, sub = lane / lpp, ppw = 32 / lpp;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float inv_c = 1.0f / (float)c;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 cacc[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) cacc[s] = zero4;
  auto group_sum = [&](float v) {
    for (int o = lpp >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  auto dot4 = [](const float4& x, const float4& y) { return x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w; };
  auto mask4 = [](float4& v, const float4& t) {
    v.x *= gs_lrelu_slope(t.x); v.y *= gs_lrelu_slope(t.y); v.z *= gs_lrelu_slope(t.z); v.w *= gs_lrelu_slope(t.w);
  };
  for (long long base = warp * ppw * R; base < rows; base += nwarps * ppw * R) {
    float4 t[R][SLOTS], d[R][SLOTS], w[R][SLOTS];
    bool ok[R];
    size_t off[R];
    float rr[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const long long row = base + (long long)k * ppw + sub;
      ok[k] = row < rows;
      off[k] = (size_t)(ok[k] ? row : 0) * c;
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        const int ch = (s * lpp + li) * 4;
        t[k][s] = ok[k] ? *reinterpret_cast<const float4*>(a + off[k] + ch) : zero4;
        if (MODE != 0) d[k][s] = ok[k] ? *reinterpret_cast<const float4*>(dy + off[k] + ch) : zero4;
        if (MODE == 2) w[k][s] = ok[k] ? *reinterpret_cast<const float4*>(u + off[k] + ch) : zero4;
      }
      if (MODE != 0) rr[k] = ok[k] ? __ldg(rin + row) : 0.0f;       // (row itself: off[k] / c would be a 64-bit division per pixel)
      if (MODE != 0 && (flags & 4)) {
        // "y form": `a` holds the NORMALISED output y = a * r of the fused conv + pixel-norm layer; a = y / r
        const float inv = ok[k] ? 1.0f / rr[k] : 0.0f;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) { t[k][s].x *= inv; t[k][s].y *= inv; t[k][s].z *= inv; t[k][s].w *= inv; }
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if (flags & 1) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) mask4((MODE == 2) ? w[k][s] : d[k][s], t[k][s]);
      }
      if (MODE == 0) {
        float ss = 0.0f;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) ss += dot4(t[k][s], t[k][s]);
        ss = group_sum(ss);
        const float r = 1.0f / sqrtf(ss * inv_c + eps);
        if (ok[k]) {
#pragma unroll
          for (int s = 0; s < SLOTS; ++s) {
            const int ch = (s * lpp + li) * 4;
            *reinterpret_cast<float4*>(out + off[k] + ch) = make_float4(t[k][s].x * r, t[k][s].y * r, t[k][s].z * r, t[k][s].w * r);
          }
          if (li == 0) rout[base + (long long)k * ppw + sub] = r;
        }
      } else if (MODE == 1 || MODE == 3) {
        float dot = 0.0f;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) dot += dot4(t[k][s], d[k][s]);
        dot = group_sum(dot);
        const float r = rr[k];
        const float kk = r * r * r * inv_c * dot;
        if (ok[k]) {
#pragma unroll
          for (int s = 0; s < SLOTS; ++s) {
            const int ch = (s * lpp + li) * 4;
            const float4 tt = t[k][s], dd = d[k][s];
            float4 o = make_float4(r * dd.x - kk * tt.x, r * dd.y - kk * tt.y, r * dd.z - kk * tt.z, r * dd.w - kk * tt.w);
            if (MODE == 3) {
              mask4(o, tt);
              cacc[s].x += o.x; cacc[s].y += o.y; cacc[s].z += o.z; cacc[s].w += o.w;
            }
            *reinterpret_cast<float4*>(out + off[k] + ch) = o;
          }
        }
      } else {
        float ud = 0.0f, ad = 0.0f, ua = 0.0f;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
          ud += dot4(w[k][s], d[k][s]);
          ad += dot4(t[k][s], d[k][s]);
          ua += dot4(w[k][s], t[k][s]);
        }
        ud = group_sum(ud); ad = group_sum(ad); ua = group_sum(ua);
        const float r = rr[k];
        const float r3c = r * r * r * inv_c;
        const float ka = -r3c * ud + 3.0f * r3c * r * r * inv_c * ua * ad;
        const float ku = -r3c * ad, kd = -r3c * ua;
        if (ok[k]) {
#pragma unroll
          for (int s = 0; s < SLOTS; ++s) {
            const int ch = (s * lpp + li) * 4;
            const float4 tt = t[k][s], dd = d[k][s], ww = w[k][s];
            float4 o = make_float4(ka * tt.x + ku * ww.x + kd * dd.x, ka * tt.y + ku * ww.y + kd * dd.y,
                                   ka * tt.z + ku * ww.z + kd * dd.z, ka * tt.w + ku * ww.w + kd * dd.w);
            if (flags & 2) mask4(o, tt);
            *reinterpret_cast<float4*>(out + off[k] + ch) = o;
            if (flags & 8) {
              // second output: the MODE 1 result for the same (masked) u, r * u - r^3 / c * <a, u> * a; <a, u> = ua
              const float k1 = r3c * ua;
              *reinterpret_cast<float4*>(rout + off[k] + ch) =
                  make_float4(r * ww.x - k1 * tt.x, r * ww.y - k1 * tt.y, r * ww.z - k1 * tt.z, r * ww.w - k1 * tt.w);
            }
          }
        }
      }
    }
  }
  if (MODE == 3 && rout != nullptr) {
    // rout doubles as the bias-gradient output [c] in this mode
    __shared__ float cs[512];
    for (int i = threadIdx.x; i < c; i += blockDim.x) cs[i] = 0.0f;
    __syncthreads();
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int ch = (s * lpp + li) * 4;
      atomicAdd(&cs[ch], cacc[s].x); atomicAdd(&cs[ch + 1], cacc[s].y);
      atomicAdd(&cs[ch + 2], cacc[s].z); atomicAdd(&cs[ch + 3], cacc[s].w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < c; i += blockDim.x) atomicAdd(rout + i, cs[i]);
  }
}

template <int MODE>
static void pn_vec_launch(const float* a, const float* r, const float* dy, const float* u, float* out, float* rout, long long rows,
                          int c, float eps, int lpp, int flags, cudaStream_t st) {
  const int slots = (c / 4) / lpp;
  const int R = slots == 1 ? 4 : slots == 2 ? 2 : 1;
  const long long ppw = 32 / lpp, warps_per_block = EW_BLOCK / 32;
  long long b = (rows + warps_per_block * ppw * R - 1) / (warps_per_block * ppw * R);
  const long long cap = (long long)gs_num_sms() * 16;
  if (b < 1) b = 1;
  const int blocks = (int)(b < cap ? b : cap);
  if (slots == 1) pixel_norm_vec_kernel<MODE, 1, 4><<<blocks, EW_BLOCK, 0, st>>>(a, r, dy, u, out, rout, rows, c, eps, lpp, flags);
  else if (slots == 2) pixel_norm_vec_kernel<MODE, 2, 2><<<blocks, EW_BLOCK, 0, st>>>(a, r, dy, u, out, rout, rows, c, eps, lpp, flags);
  else pixel_norm_vec_kernel<MODE, 4, 1><<<blocks, EW_BLOCK, 0, st>>>(a, r, dy, u, out, rout, rows, c, eps, lpp, flags);
}

// lanes per pixel of the vectorised kernel, or 0 when c is not covered
static int pn_lpp(int c) {
  if (c % 4) return 0;
  const int v = c / 4;
  if (v <= 32) return (v & (v - 1)) == 0 ? v : 0;
  return (v == 64 || v == 128) ? 32 : 0;     // 2 or 4 float4 per lane
}
static int pn_grid(long long rows) {
  long long warps_per_block = EW_BLOCK / 32;
  long long b = (rows + warps_per_block * 4 - 1) / (warps_per_block * 4);
  long long cap = (long long)gs_num_sms() * 16;
  if (b < 1) b = 1;
  return (int)(b < cap ? b : cap);
}
extern "C" int gs_pixel_norm_fwd(const float* a, float* y, float* r, long long rows, int c, float eps, void* stream) {
  GS_CHECK_ARG(rows >= 0 && c > 0, "pixel_norm_fwd: bad shape");
  if (rows == 0) return GS_OK;
  if (const int lpp = pn_lpp(c)) pn_vec_launch<0>(a, nullptr, nullptr, nullptr, y, r, rows, c, eps, lpp, 0, ST);
  else pixel_norm_kernel<0><<<pn_grid(rows), EW_BLOCK, 0, ST>>>(a, nullptr, nullptr, nullptr, y, r, rows, c, eps);
  GS_CHECK_LAUNCH("pixel_norm_fwd");
  return GS_OK;
}
extern "C" int gs_pixel_norm_bwd(const float* a, const float* r, const float* dy, float* da, long long rows, int c, void* stream) {
  GS_CHECK_ARG(rows >= 0 && c > 0, "pixel_norm_bwd: bad shape");
  if (rows == 0) return GS_OK;
  if (const int lpp = pn_lpp(c)) pn_vec_launch<1>(a, r, dy, nullptr, da, nullptr, rows, c, 0.f, lpp, 0, ST);
  else pixel_norm_kernel<1><<<pn_grid(rows), EW_BLOCK, 0, ST>>>(a, r, dy, nullptr, da, nullptr, rows, c, 0.f);
  GS_CHECK_LAUNCH("pixel_norm_bwd");
  return GS_OK;
}
extern "C" int gs_pixel_norm_bwd2(const float* a, const float* r, const float* dy, const float* u, float* ga, long long rows,
                                  int c, void* stream) {
  GS_CHECK_ARG(rows >= 0 && c > 0, "pixel_norm_bwd2: bad shape");
  if (rows == 0) return GS_OK;
  if (const int lpp = pn_lpp(c)) pn_vec_launch<2>(a, r, dy, u, ga, nullptr, rows, c, 0.f, lpp, 0, ST);
  else pixel_norm_kernel<2><<<pn_grid(rows), EW_BLOCK, 0, ST>>>(a, r, dy, u, ga, nullptr, rows, c, 0.f);
  GS_CHECK_LAUNCH("pixel_norm_bwd2");
  return GS_OK;
}

extern "C" int gs_pixel_norm_bwd_mask(const float* a, const float* r, const float* dy, float* dz, float* colsum, long long rows,
                                      int c, void* stream) {
  GS_CHECK_ARG(rows >= 0 && c > 0 && c % 4 == 0 && c <= 256, "pixel_norm_bwd_mask: needs c %% 4 == 0, c <= 256 (got %d)", c);
  if (colsum) GS_CUDA(cudaMemsetAsync(colsum, 0, (size_t)c * sizeof(float), ST));
  if (rows == 0) return GS_OK;
  if (const int lpp = pn_lpp(c)) pn_vec_launch<3>(a, r, dy, nullptr, dz, colsum, rows, c, 0.f, lpp, 0, ST);
  else pixel_norm_kernel<3><<<pn_grid(rows), EW_BLOCK, 0, ST>>>(a, r, dy, nullptr, dz, colsum, rows, c, 0.f);
  GS_CHECK_LAUNCH("pixel_norm_bwd_mask");
  return GS_OK;
}
extern "C" int gs_lrelu_mask_mul_colsum(const float* v, const float* y, float* out, float* colsum, long long rows, int c,
                                        void* stream) {
  GS_CHECK_ARG(rows >= 0 && c > 0 && c % 4 == 0 && c <= 256 && 256 % (c / 4) == 0,
               "lrelu_mask_mul_colsum: needs c %% 4 == 0, c <= 256, c / 4 a divisor of 256 (got %d)", c);
  GS_CUDA(cudaMemsetAsync(colsum, 0, (size_t)c * sizeof(float), ST));
  if (rows == 0) return GS_OK;
  const size_t n4 = (size_t)rows * c / 4;
  mask_mul_colsum_kernel<<<ew_grid(n4 * 4), EW_BLOCK, 0, ST>>>(reinterpret_cast<const float4*>(v), reinterpret_cast<const float4*>(y),
                                                               reinterpret_cast<float4*>(out), colsum, n4, c / 4);
  GS_CHECK_LAUNCH("lrelu_mask_mul_colsum");
  return GS_OK;
}

extern "C" int gs_pixel_norm_bwd_premask(const float* a, const float* r, const float* u, float* out, long long rows, int c,
                                         void* stream) {
  const int lpp = pn_lpp(c);
  GS_CHECK_ARG(rows >= 0 && c > 0 && lpp > 0, "pixel_norm_bwd_premask: unsupported channel count %d", c);
  if (rows == 0) return GS_OK;
  pn_vec_launch<1>(a, r, u, nullptr, out, nullptr, rows, c, 0.f, lpp, 1, ST);
  GS_CHECK_LAUNCH("pixel_norm_bwd_premask");
  return GS_OK;
}
extern "C" int gs_pixel_norm_bwd2_masked(const float* a, const float* r, const float* dy, const float* u, float* ga,
                                         long long rows, int c, void* stream) {
  const int lpp = pn_lpp(c);
  GS_CHECK_ARG(rows >= 0 && c > 0 && lpp > 0, "pixel_norm_bwd2_masked: unsupported channel count %d", c);
  if (rows == 0) return GS_OK;
  pn_vec_launch<2>(a, r, dy, u, ga, nullptr, rows, c, 0.f, lpp, 3, ST);
  GS_CHECK_LAUNCH("pixel_norm_bwd2_masked");
  return GS_OK;
}

// "y form" of the three fused pixel-norm / leaky-relu gradient kernels: the layer kept only its normalised output
// y = a * r (conv epilogue GS_EPI_PIXEL_NORM) and r; a = y / r is rebuilt in registers.
extern "C" int gs_pixel_norm_bwd_mask_y(const float* y, const float* r, const float* dy, float* dz, float* colsum, long long rows,
                                        int c, void* stream) {
  const int lpp = pn_lpp(c);
  GS_CHECK_ARG(rows >= 0 && c > 0 && lpp > 0 && c <= 512, "pixel_norm_bwd_mask_y: unsupported channel count %d", c);
  if (colsum) GS_CUDA(cudaMemsetAsync(colsum, 0, (size_t)c * sizeof(float), ST));
  if (rows == 0) return GS_OK;
  pn_vec_launch<3>(y, r, dy, nullptr, dz, colsum, rows, c, 0.f, lpp, 4, ST);
  GS_CHECK_LAUNCH("pixel_norm_bwd_mask_y");
  return GS_OK;
}
extern "C" int gs_pixel_norm_bwd_premask_y(const float* y, const float* r, const float* u, float* out, long long rows, int c,
                                           void* stream) {
  const int lpp = pn_lpp(c);
  GS_CHECK_ARG(rows >= 0 && c > 0 && lpp > 0, "pixel_norm_bwd_premask_y: unsupported channel count %d", c);
  if (rows == 0) return GS_OK;
  pn_vec_launch<1>(y, r, u, nullptr, out, nullptr, rows, c, 0.f, lpp, 1 | 4, ST);
  GS_CHECK_LAUNCH("pixel_norm_bwd_premask_y");
  return GS_OK;
}
extern "C" int gs_pixel_norm_bwd2_masked_y(const float* y, const float* r, const float* dy, const float* u, float* ga,
                                           long long rows, int c, void* stream) {
  const int lpp = pn_lpp(c);
  GS_CHECK_ARG(rows >= 0 && c > 0 && lpp > 0, "pixel_norm_bwd2_masked_y: unsupported channel count %d", c);
  if (rows == 0) return GS_OK;
  pn_vec_launch<2>(y, r, dy, u, ga, nullptr, rows, c, 0.f, lpp, 3 | 4, ST);
  GS_CHECK_LAUNCH("pixel_norm_bwd2_masked_y");
  return GS_OK;
}

// Both second-order pieces of the fused pixel-norm / leaky-relu backward from ONE pass over (y, r, dy, u):
// ga = gs_pixel_norm_bwd2_masked_y(...), gdy = gs_pixel_norm_bwd_premask_y(y, r, u) -- 20 instead of 32 bytes per element.
extern "C" int gs_pixel_norm_bwd2_pair_y(const float* y, const float* r, const float* dy, const float* u, float* ga, float* gdy,
                                         long long rows, int c, void* stream) {
  const int lpp = pn_lpp(c);
  GS_CHECK_ARG(rows >= 0 && c > 0 && lpp > 0, "pixel_norm_bwd2_pair_y: unsupported channel count %d", c);
  if (rows == 0) return GS_OK;
  pn_vec_launch<2>(y, r, dy, u, ga, gdy, rows, c, 0.f, lpp, 3 | 4 | 8, ST);
  GS_CHECK_LAUNCH("pixel_norm_bwd2_pair_y");
  return GS_OK;
}

extern "C" int gs_upscale2d(const float* in, float* out, int n, int h, int w, int c, int fh, int fw, float scale, void* stream) {
  GS_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && fh > 0 && fw > 0, "upscale2d: bad shape");
  size_t total = (size_t)n * h * fh * w * fw * c;
  upscale_kernel<<<ew_grid(total), EW_BLOCK, 0, ST>>>(in, out, n, h, w, c, fh, fw, scale);
  GS_CHECK_LAUNCH("upscale2d");
  return GS_OK;
}
extern "C" int gs_pool2d(const float* in, float* out, int n, int h, int w, int c, int fh, int fw, float scale, void* stream) {
  GS_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && fh > 0 && fw > 0, "pool2d: bad shape");
  size_t total = (size_t)n * h * w * c;
  pool_kernel<<<ew_grid(total, 1), EW_BLOCK, 0, ST>>>(in, out, n, h, w, c, fh, fw, scale);
  GS_CHECK_LAUNCH("pool2d");
  return GS_OK;
}
extern "C" int gs_transpose_inner(const float* in, float* out, int n, int a, int b, void* stream) {
  GS_CHECK_ARG(n > 0 && a > 0 && b > 0, "transpose_inner: bad shape");
  size_t total = (size_t)n * a * b;
  transpose_inner_kernel<<<ew_grid(total), EW_BLOCK, 0, ST>>>(in, out, n, a, b);
  GS_CHECK_LAUNCH("transpose_inner");
  return GS_OK;
}
extern "C" int gs_row_dot(const float* a, const float* b, float* out, int rows, long long e, void* stream) {
  GS_CHECK_ARG(rows > 0 && e > 0, "row_dot: bad shape");
  GS_CUDA(cudaMemsetAsync(out, 0, (size_t)rows * sizeof(float), ST));
  int gx = gs_cdiv(e, EW_BLOCK * 8);
  int cap = gs_num_sms() * 8 / rows + 1;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, (unsigned)rows);
  row_dot_kernel<<<grid, EW_BLOCK, 0, ST>>>(a, b, out, (size_t)e);
  GS_CHECK_LAUNCH("row_dot");
  return GS_OK;
}
extern "C" int gs_row_scale(const float* a, const float* s, float* out, int rows, long long e, float alpha, void* stream) {
  GS_CHECK_ARG(rows > 0 && e > 0, "row_scale: bad shape");
  int gx = gs_cdiv(e, EW_BLOCK * 4);
  int cap = gs_num_sms() * 16 / rows + 1;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, (unsigned)rows);
  row_scale_kernel<<<grid, EW_BLOCK, 0, ST>>>(a, s, out, (size_t)e, alpha);
  GS_CHECK_LAUNCH("row_scale");
  return GS_OK;
}

static int stddev_launch(int mode, const float* x, const float* df, const float* u, float* out, float* stat, int b,
                         long long e, int groups, float eps, cudaStream_t st) {
  GS_CHECK_ARG(b > 0 && e > 0 && groups > 0 && b % groups == 0, "batch_stddev: batch %d not divisible by groups %d", b, groups);
  int m = b / groups;
  if (mode != 1) GS_CUDA(cudaMemsetAsync(stat, 0, (size_t)m * sizeof(float), st));
  int gx = gs_cdiv(e, EW_BLOCK);
  if (gx > 64) gx = 64;
  dim3 grid((unsigned)gx, (unsigned)m);
  if (mode == 0) stddev_kernel<0><<<grid, EW_BLOCK, 0, st>>>(x, df, u, out, stat, groups, m, (size_t)e, eps);
  else if (mode == 1) stddev_kernel<1><<<grid, EW_BLOCK, 0, st>>>(x, df, u, out, stat, groups, m, (size_t)e, eps);
  else stddev_kernel<2><<<grid, EW_BLOCK, 0, st>>>(x, df, u, out, stat, groups, m, (size_t)e, eps);
  GS_CHECK_LAUNCH("batch_stddev");
  return GS_OK;
}
extern "C" int gs_batch_stddev_fwd(const float* x, float* stat, int b, long long e, int groups, float eps, void* stream) {
  return stddev_launch(0, x, nullptr, nullptr, nullptr, stat, b, e, groups, eps, ST);
}
extern "C" int gs_batch_stddev_bwd(const float* x, const float* df, float* dx, int b, long long e, int groups, float eps, void* stream) {
  return stddev_launch(1, x, df, nullptr, dx, nullptr, b, e, groups, eps, ST);
}
extern "C" int gs_batch_stddev_bwd2(const float* x, const float* df, const float* u, float* gx, float* q, int b, long long e,
                                    int groups, float eps, void* stream) {
  return stddev_launch(2, x, df, u, gx, q, b, e, groups, eps, ST);
}

extern "C" int gs_embedding_fwd(const float* table, const long long* idx, float* out, int b, int units, float alpha, void* stream) {
  GS_CHECK_ARG(b > 0 && units > 0, "embedding_fwd: bad shape");
  embedding_fwd_kernel<<<ew_grid((size_t)b * units, 1), EW_BLOCK, 0, ST>>>(table, idx, out, b, units, alpha);
  GS_CHECK_LAUNCH("embedding_fwd");
  return GS_OK;
}
extern "C" int gs_embedding_bwd(const float* dy, const long long* idx, float* dtable, int b, int rows, int units, float alpha,
                                void* stream) {
  GS_CHECK_ARG(b > 0 && units > 0 && rows > 0, "embedding_bwd: bad shape");
  GS_CUDA(cudaMemsetAsync(dtable, 0, (size_t)rows * units * sizeof(float), ST));
  embedding_bwd_kernel<<<ew_grid((size_t)b * units, 1), EW_BLOCK, 0, ST>>>(dy, idx, dtable, b, units, alpha);
  GS_CHECK_LAUNCH("embedding_bwd");
  return GS_OK;
}

extern "C" int gs_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                            float eps, long long t, float grad_scale, void* stream) {
  GS_CHECK_ARG(n >= 0 && t >= 1, "adam_step: need n >= 0 and step t >= 1");
  if (n == 0) return GS_OK;
  double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)beta1, (double)t));
  adam_kernel<<<ew_grid(n), EW_BLOCK, 0, ST>>>(p, g, m, v, (size_t)n, (float)lr_t, beta1, beta2, eps, grad_scale);
  GS_CHECK_LAUNCH("adam_step");
  return GS_OK;
}

// The slice [*lo, *hi) of an n-element flat buffer that rank `rank` of `world` owns in gs_adam_step_allreduce
// (multiples of 4 elements: 16-byte multimem accesses).
extern "C" int gs_adam_slice(long long n, int rank, int world, long long* lo, long long* hi) {
  GS_CHECK_ARG(n >= 0 && n % 4 == 0 && world >= 1 && rank >= 0 && rank < world && lo && hi, "adam_slice: bad arguments");
  const long long quads = n / 4, per = (quads + world - 1) / world;
  const long long a = per * rank < quads ? per * rank : quads, b = per * (rank + 1) < quads ? per * (rank + 1) : quads;
  *lo = 4 * a;
  *hi = 4 * b;
  return GS_OK;
}

extern "C" int gs_adam_step_allreduce(const float* p_local, float* m, float* v, const float* grad_multicast,
                                      float* param_multicast, const float* const* grad_peers, float* const* param_peers,
                                      long long n, int rank, int world, float lr, float beta1, float beta2, float eps,
                                      long long t, float grad_scale, void* stream) {
  GS_CHECK_ARG(n >= 0 && n % 4 == 0 && t >= 1 && world >= 1 && rank >= 0 && rank < world, "adam_step_allreduce: bad arguments");
  const bool mc = grad_multicast != nullptr && param_multicast != nullptr;
  GS_CHECK_ARG(mc || (grad_peers != nullptr && param_peers != nullptr),
               "adam_step_allreduce: needs multicast addresses or the arrays of peer pointers");
  long long lo = 0, hi = 0;
  int rc = gs_adam_slice(n, rank, world, &lo, &hi);
  if (rc) return rc;
  if (hi <= lo) return GS_OK;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)beta1, (double)t));
  const long long quads = (hi - lo) / 4;
  long long blocks = (quads + EW_BLOCK - 1) / EW_BLOCK;
  const long long cap = (long long)gs_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (mc)
    adam_reduce_kernel<0><<<(unsigned)blocks, EW_BLOCK, 0, ST>>>(p_local, m, v, grad_multicast, param_multicast, nullptr, nullptr, world,
                                                                  lo / 4, hi / 4, (float)lr_t, beta1, beta2, eps, grad_scale);
  else
    adam_reduce_kernel<1><<<(unsigned)blocks, EW_BLOCK, 0, ST>>>(p_local, m, v, nullptr, nullptr, grad_peers, param_peers, world, lo / 4,
                                                                  hi / 4, (float)lr_t, beta1, beta2, eps, grad_scale);
  GS_CHECK_LAUNCH("adam_step_allreduce");
  return GS_OK;
}
