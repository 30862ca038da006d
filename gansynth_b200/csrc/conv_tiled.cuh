// Shared-memory tiled fp32 (FFMA) NHWC 3x3 convolution kernels for sm_100a:
//   conv_c_tiled_kernel<S,TN>  gather form (forward of conv2d, stride 1|2; with flip = dgrad of stride 1)
//   conv_t2_tiled_kernel<TN>   transposed stride-2 form (conv2d_transpose forward / dgrad of stride 2)
//   conv_w_tiled_kernel<S,TC>  weight-gradient form
// Channel counts must be multiples of 4 (float4 paths); everything else goes to conv_naive.cuh.
// These are the exact-fp32 path; the tcgen05 kernels (conv_tc.cuh) are checked against them.
#pragma once
#include "conv_naive.cuh"

// ------------------------------------------------------------------------------------------------
// B tile loader: Bs[tap][k][n] (k = 0..7 of this K chunk, n = 0..TN-1) from the weight tensor.
// w_is_kn: memory is [tap][K][N] (n contiguous); else [tap][N][K] (k contiguous).
// flip: logical tap t reads memory tap 8 - t (180-degree rotated filter).
template <int TN, int NT>
__device__ __forceinline__ void gs_load_b_tile(float* __restrict__ Bs, const float* __restrict__ w, int kdim,
                                               int ndim, int k0, int n0, int w_is_kn, int flip, int tid) {
  if (w_is_kn) {
    constexpr int N4 = TN / 4;
    for (int idx = tid; idx < 9 * 8 * N4; idx += NT) {
      int c4 = idx % N4;
      int k = (idx / N4) % 8;
      int tap = idx / (N4 * 8);
      int st = flip ? 8 - tap : tap;
      int kk = k0 + k, nn = n0 + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kk < kdim && nn < ndim) v = *reinterpret_cast<const float4*>(w + ((size_t)st * kdim + kk) * ndim + nn);
      *reinterpret_cast<float4*>(Bs + (tap * 8 + k) * TN + c4 * 4) = v;
    }
  } else {
    for (int idx = tid; idx < 9 * TN * 2; idx += NT) {
      int nl = idx % TN;
      int q = (idx / TN) & 1;
      int tap = idx / (TN * 2);
      int st = flip ? 8 - tap : tap;
      int kk = k0 + q * 4, nn = n0 + nl;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kk < kdim && nn < ndim) v = *reinterpret_cast<const float4*>(w + ((size_t)st * ndim + nn) * kdim + kk);
      float* b = Bs + (tap * 8 + q * 4) * TN + nl;
      b[0] = v.x;
      b[TN] = v.y;
      b[2 * TN] = v.z;
      b[3 * TN] = v.w;
    }
  }
}

// ------------------------------------------------------------------------------------------------
template <int S, int TN>
struct CTile {
  static constexpr int NT = 128;
  static constexpr int NCG = TN / 8;
  static constexpr int NPG = NT / NCG;
  static constexpr int TW = 16, TH = NPG / 2;
  static constexpr int HR = (TH - 1) * S + 3;
  static constexpr int HC = (TW - 1) * S + 3;
  static constexpr int PITCH = (S == 1) ? 20 : 40;
  static constexpr int KS0 = HR * PITCH;
  static constexpr int KS = KS0 + ((12 - (KS0 % 8)) % 8);  // KS % 8 == 4: conflict-free transposed stores
  static constexpr int AS = 8 * KS;
  static constexpr int BS = 9 * 8 * TN;
  static constexpr int SMEM = (AS + BS) * 4;
};

// x [n,h,w,kdim] (NHWC), y [n,oh,ow,ndim].  kdim/ndim are the contraction / output channel counts.
template <int S, int TN>
__global__ void __launch_bounds__(128)
conv_c_tiled_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                    float* __restrict__ y, int n_img, int h, int wd, int kdim, int ndim, int oh, int ow, int pb,
                    int w_is_kn, int flip, float alpha, int act, int tiles_h, int tiles_w) {
  using T = CTile<S, TN>;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = smem + T::AS;
  const int tid = threadIdx.x;
  const int cg = tid % T::NCG, pg = tid / T::NCG;
  const int prow = pg >> 1, pcol0 = (pg & 1) * 8;
  int t = blockIdx.x;
  const int tw_ = t % tiles_w;
  t /= tiles_w;
  const int th_ = t % tiles_h;
  const int n = t / tiles_h;
  const int oh0 = th_ * T::TH, ow0 = tw_ * T::TW;
  const int n0 = blockIdx.y * TN;
  const int ih0 = oh0 * S - pb, iw0 = ow0 * S - pb;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < kdim; k0 += 8) {
    for (int idx = tid; idx < T::HR * T::HC * 2; idx += T::NT) {
      int q = idx & 1, p = idx >> 1;
      int hr = p / T::HC, hc = p % T::HC;
      int ih = ih0 + hr, iw = iw0 + hc;
      int c = k0 + q * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ih >= 0 && ih < h && iw >= 0 && iw < wd && c < kdim)
        v = *reinterpret_cast<const float4*>(x + (((size_t)n * h + ih) * wd + iw) * kdim + c);
      int off = (S == 1) ? hr * 20 + hc : hr * 40 + (hc & 1) * 20 + (hc >> 1);
      float* a = As + (q * 4) * T::KS + off;
      a[0] = v.x;
      a[T::KS] = v.y;
      a[2 * T::KS] = v.z;
      a[3 * T::KS] = v.w;
    }
    gs_load_b_tile<TN, T::NT>(Bs, w, kdim, ndim, k0, n0, w_is_kn, flip, tid);
    __syncthreads();
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const float* arow = As + k * T::KS + (prow * S + kh) * T::PITCH + pcol0;
        float a0[12];
        float a1[8];
        {
          float4 v0 = *reinterpret_cast<const float4*>(arow);
          float4 v1 = *reinterpret_cast<const float4*>(arow + 4);
          float4 v2 = *reinterpret_cast<const float4*>(arow + 8);
          a0[0] = v0.x; a0[1] = v0.y; a0[2] = v0.z; a0[3] = v0.w;
          a0[4] = v1.x; a0[5] = v1.y; a0[6] = v1.z; a0[7] = v1.w;
          a0[8] = v2.x; a0[9] = v2.y; a0[10] = v2.z; a0[11] = v2.w;
        }
        if (S == 2) {
          float4 v0 = *reinterpret_cast<const float4*>(arow + 20);
          float4 v1 = *reinterpret_cast<const float4*>(arow + 24);
          a1[0] = v0.x; a1[1] = v0.y; a1[2] = v0.z; a1[3] = v0.w;
          a1[4] = v1.x; a1[5] = v1.y; a1[6] = v1.z; a1[7] = v1.w;
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float* b = Bs + ((kh * 3 + kw) * 8 + k) * TN + cg * 4;
          float4 b0 = *reinterpret_cast<const float4*>(b);
          float4 b1 = *reinterpret_cast<const float4*>(b + TN / 2);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float a;
            if (S == 1) a = a0[i + kw];
            else a = (kw == 1) ? a1[i] : a0[i + (kw >> 1)];
            acc[i][0] = fmaf(a, b0.x, acc[i][0]);
            acc[i][1] = fmaf(a, b0.y, acc[i][1]);
            acc[i][2] = fmaf(a, b0.z, acc[i][2]);
            acc[i][3] = fmaf(a, b0.w, acc[i][3]);
            acc[i][4] = fmaf(a, b1.x, acc[i][4]);
            acc[i][5] = fmaf(a, b1.y, acc[i][5]);
            acc[i][6] = fmaf(a, b1.z, acc[i][6]);
            acc[i][7] = fmaf(a, b1.w, acc[i][7]);
          }
        }
      }
    }
    __syncthreads();
  }

  const int oy = oh0 + prow;
  if (oy >= oh) return;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    int co = n0 + half * (TN / 2) + cg * 4;
    if (co >= ndim) continue;
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bv = *reinterpret_cast<const float4*>(bias + co);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int ox = ow0 + pcol0 + i;
      if (ox >= ow) continue;
      float4 o;
      o.x = fmaf(acc[i][half * 4 + 0], alpha, bv.x);
      o.y = fmaf(acc[i][half * 4 + 1], alpha, bv.y);
      o.z = fmaf(acc[i][half * 4 + 2], alpha, bv.z);
      o.w = fmaf(acc[i][half * 4 + 3], alpha, bv.w);
      if (act == 1) { o.x = gs_lrelu(o.x); o.y = gs_lrelu(o.y); o.z = gs_lrelu(o.z); o.w = gs_lrelu(o.w); }
      *reinterpret_cast<float4*>(y + (((size_t)n * oh + oy) * ow + ox) * ndim + co) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Transposed stride-2 3x3 (pad_before 0): out[n, 2i+ph, 2j+pw, f] = alpha * sum_k
//   ph=0: in[i][..]*W(0,.) + in[i-1][..]*W(2,.)   ph=1: in[i][..]*W(1,.)      (same along columns)
// in [n,ih,iw,kdim] (small side), out [n,2ih,2iw,ndim].
template <int TN>
struct T2Tile {
  static constexpr int NT = 128;
  static constexpr int NCG = TN / 4;
  static constexpr int NPG = NT / NCG;
  static constexpr int TW = 16, TH = NPG / 2;
  static constexpr int HR = TH + 1;
  static constexpr int KS0 = HR * 20;
  static constexpr int KS = KS0 + ((12 - (KS0 % 8)) % 8);
  static constexpr int AS = 8 * KS;
  static constexpr int BS = 9 * 8 * TN;
  static constexpr int SMEM = (AS + BS) * 4;
};

template <int TN>
__global__ void __launch_bounds__(128)
conv_t2_tiled_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                     float* __restrict__ out, int n_img, int ih, int iw, int kdim, int ndim, int w_is_kn,
                     float alpha, int act, int tiles_h, int tiles_w) {
  using T = T2Tile<TN>;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = smem + T::AS;
  const int tid = threadIdx.x;
  const int cg = tid % T::NCG, pg = tid / T::NCG;
  const int prow = pg >> 1, pcol0 = (pg & 1) * 8;
  int t = blockIdx.x;
  const int tw_ = t % tiles_w;
  t /= tiles_w;
  const int th_ = t % tiles_h;
  const int n = t / tiles_h;
  const int i0 = th_ * T::TH, j0 = tw_ * T::TW;
  const int n0 = blockIdx.y * TN;

  float acc[8][4][4];  // [position][phase][channel]
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][p][j] = 0.f;

  for (int k0 = 0; k0 < kdim; k0 += 8) {
    for (int idx = tid; idx < T::HR * 17 * 2; idx += T::NT) {
      int q = idx & 1, p = idx >> 1;
      int hr = p / 17, hc = p % 17;
      int r = i0 - 1 + hr, c = j0 - 1 + hc;
      int ch = k0 + q * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r >= 0 && r < ih && c >= 0 && c < iw && ch < kdim)
        v = *reinterpret_cast<const float4*>(in + (((size_t)n * ih + r) * iw + c) * kdim + ch);
      float* a = As + (q * 4) * T::KS + hr * 20 + hc;
      a[0] = v.x;
      a[T::KS] = v.y;
      a[2 * T::KS] = v.z;
      a[3 * T::KS] = v.w;
    }
    gs_load_b_tile<TN, T::NT>(Bs, w, kdim, ndim, k0, n0, w_is_kn, 0, tid);
    __syncthreads();
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
      float up[12], cu[12];
      {
        const float* r0 = As + k * T::KS + prow * 20 + pcol0;
        float4 v0 = *reinterpret_cast<const float4*>(r0);
        float4 v1 = *reinterpret_cast<const float4*>(r0 + 4);
        float4 v2 = *reinterpret_cast<const float4*>(r0 + 8);
        up[0] = v0.x; up[1] = v0.y; up[2] = v0.z; up[3] = v0.w;
        up[4] = v1.x; up[5] = v1.y; up[6] = v1.z; up[7] = v1.w;
        up[8] = v2.x; up[9] = v2.y; up[10] = v2.z; up[11] = v2.w;
        const float* r1 = r0 + 20;
        v0 = *reinterpret_cast<const float4*>(r1);
        v1 = *reinterpret_cast<const float4*>(r1 + 4);
        v2 = *reinterpret_cast<const float4*>(r1 + 8);
        cu[0] = v0.x; cu[1] = v0.y; cu[2] = v0.z; cu[3] = v0.w;
        cu[4] = v1.x; cu[5] = v1.y; cu[6] = v1.z; cu[7] = v1.w;
        cu[8] = v2.x; cu[9] = v2.y; cu[10] = v2.z; cu[11] = v2.w;
      }
      float4 wv[9];
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) wv[tp] = *reinterpret_cast<const float4*>(Bs + (tp * 8 + k) * TN + cg * 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        // local halo col of position j is i+1, of j-1 is i
        const float c1 = cu[i + 1], c0 = cu[i], u1 = up[i + 1], u0 = up[i];
#define GS_FMA4(dst, a, wq)                    \
  dst[0] = fmaf(a, wq.x, dst[0]);              \
  dst[1] = fmaf(a, wq.y, dst[1]);              \
  dst[2] = fmaf(a, wq.z, dst[2]);              \
  dst[3] = fmaf(a, wq.w, dst[3]);
        // phase (0,0): taps (0,0)->cur[j], (0,2)->cur[j-1], (2,0)->up[j], (2,2)->up[j-1]
        GS_FMA4(acc[i][0], c1, wv[0]) GS_FMA4(acc[i][0], c0, wv[2]) GS_FMA4(acc[i][0], u1, wv[6]) GS_FMA4(acc[i][0], u0, wv[8])
        // phase (0,1): taps (0,1)->cur[j], (2,1)->up[j]
        GS_FMA4(acc[i][1], c1, wv[1]) GS_FMA4(acc[i][1], u1, wv[7])
        // phase (1,0): taps (1,0)->cur[j], (1,2)->cur[j-1]
        GS_FMA4(acc[i][2], c1, wv[3]) GS_FMA4(acc[i][2], c0, wv[5])
        // phase (1,1): tap (1,1)->cur[j]
        GS_FMA4(acc[i][3], c1, wv[4])
#undef GS_FMA4
      }
    }
    __syncthreads();
  }

  const int r = i0 + prow;
  const int co = n0 + cg * 4;
  if (r >= ih || co >= ndim) return;
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) bv = *reinterpret_cast<const float4*>(bias + co);
  const int oh = 2 * ih, ow = 2 * iw;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int c = j0 + pcol0 + i;
    if (c >= iw) continue;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int oy = 2 * r + (p >> 1), ox = 2 * c + (p & 1);
      float4 o;
      o.x = fmaf(acc[i][p][0], alpha, bv.x);
      o.y = fmaf(acc[i][p][1], alpha, bv.y);
      o.z = fmaf(acc[i][p][2], alpha, bv.z);
      o.w = fmaf(acc[i][p][3], alpha, bv.w);
      if (act == 1) { o.x = gs_lrelu(o.x); o.y = gs_lrelu(o.y); o.z = gs_lrelu(o.z); o.w = gs_lrelu(o.w); }
      *reinterpret_cast<float4*>(out + (((size_t)n * oh + oy) * ow + ox) * ndim + co) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient: dw(tap, a, b) += alpha * sum_pixels big[n, oy*S+kh-pb, ox*S+kw-pb, a] * small[n, oy, ox, b]
// out_ab: 1 -> dw[tap][a][b], 0 -> dw[tap][b][a].  dw must be zeroed; partial sums are atomically added.
template <int S, int TC>
struct WTile {
  static constexpr int NT = (TC / 4) * (TC / 4);
  static constexpr int TW = 16, TH = (S == 1) ? 8 : 4;
  static constexpr int HR = (TH - 1) * S + 3;
  static constexpr int HC = (TW - 1) * S + 3;
  static constexpr int XS = HR * HC * TC;
  static constexpr int YS = TH * TW * TC;
  static constexpr int SMEM = (XS + YS) * 4;
};

template <int S, int TC>
__global__ void __launch_bounds__((TC / 4) * (TC / 4), 1)
conv_w_tiled_kernel(const float* __restrict__ big, const float* __restrict__ small, float* __restrict__ dw,
                    int n_img, int h, int wd, int adim, int bdim, int oh, int ow, int pb, int out_ab, float alpha,
                    int tiles_h, int tiles_w) {
  using T = WTile<S, TC>;
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;
  float* Ys = smem + T::XS;
  const int tid = threadIdx.x;
  const int tb = tid % (TC / 4), ta = tid / (TC / 4);
  const int a0 = blockIdx.y * TC, b0 = blockIdx.z * TC;
  const int ntiles = n_img * tiles_h * tiles_w;

  float acc[9][4][4];
#pragma unroll
  for (int tp = 0; tp < 9; ++tp)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[tp][i][j] = 0.f;

  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    int tt = t;
    const int tw_ = tt % tiles_w;
    tt /= tiles_w;
    const int th_ = tt % tiles_h;
    const int n = tt / tiles_h;
    const int oy0 = th_ * T::TH, ox0 = tw_ * T::TW;
    const int iy0 = oy0 * S - pb, ix0 = ox0 * S - pb;
    constexpr int C4 = TC / 4;
    for (int idx = tid; idx < T::HR * T::HC * C4; idx += T::NT) {
      int c4 = idx % C4, p = idx / C4;
      int hr = p / T::HC, hc = p % T::HC;
      int iy = iy0 + hr, ix = ix0 + hc, ch = a0 + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (iy >= 0 && iy < h && ix >= 0 && ix < wd && ch < adim)
        v = *reinterpret_cast<const float4*>(big + (((size_t)n * h + iy) * wd + ix) * adim + ch);
      *reinterpret_cast<float4*>(Xs + p * TC + c4 * 4) = v;
    }
    for (int idx = tid; idx < T::TH * T::TW * C4; idx += T::NT) {
      int c4 = idx % C4, p = idx / C4;
      int r = p / T::TW, c = p % T::TW;
      int oy = oy0 + r, ox = ox0 + c, ch = b0 + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (oy < oh && ox < ow && ch < bdim)
        v = *reinterpret_cast<const float4*>(small + (((size_t)n * oh + oy) * ow + ox) * bdim + ch);
      *reinterpret_cast<float4*>(Ys + p * TC + c4 * 4) = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int p = 0; p < T::TH * T::TW; ++p) {
      const int r = p / T::TW, c = p % T::TW;
      const float4 bv = *reinterpret_cast<const float4*>(Ys + p * TC + tb * 4);
      const float* xb = Xs + ((r * S) * T::HC + c * S) * TC + ta * 4;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float4 av = *reinterpret_cast<const float4*>(xb + (kh * T::HC + kw) * TC);
          float(*d)[4] = acc[kh * 3 + kw];
          d[0][0] = fmaf(av.x, bv.x, d[0][0]); d[0][1] = fmaf(av.x, bv.y, d[0][1]);
          d[0][2] = fmaf(av.x, bv.z, d[0][2]); d[0][3] = fmaf(av.x, bv.w, d[0][3]);
          d[1][0] = fmaf(av.y, bv.x, d[1][0]); d[1][1] = fmaf(av.y, bv.y, d[1][1]);
          d[1][2] = fmaf(av.y, bv.z, d[1][2]); d[1][3] = fmaf(av.y, bv.w, d[1][3]);
          d[2][0] = fmaf(av.z, bv.x, d[2][0]); d[2][1] = fmaf(av.z, bv.y, d[2][1]);
          d[2][2] = fmaf(av.z, bv.z, d[2][2]); d[2][3] = fmaf(av.z, bv.w, d[2][3]);
          d[3][0] = fmaf(av.w, bv.x, d[3][0]); d[3][1] = fmaf(av.w, bv.y, d[3][1]);
          d[3][2] = fmaf(av.w, bv.z, d[3][2]); d[3][3] = fmaf(av.w, bv.w, d[3][3]);
        }
    }
    __syncthreads();
  }

#pragma unroll
  for (int tp = 0; tp < 9; ++tp)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int a = a0 + ta * 4 + i;
      if (a >= adim) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int b = b0 + tb * 4 + j;
        if (b >= bdim) continue;
        size_t o = out_ab ? ((size_t)tp * adim + a) * bdim + b : ((size_t)tp * bdim + b) * adim + a;
        atomicAdd(dw + o, acc[tp][i][j] * alpha);
      }
    }
}
