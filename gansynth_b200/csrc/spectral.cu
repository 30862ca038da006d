// Spectral front-end kernels (reference spectral_ops.py:45-94 convert_to_spectrogram and :97-149
// convert_to_waveform) for the reference configuration: frame 2048, hop 512 (75 % overlap),
// 1024 bins after the DC drop.  HBM-bound: one pass over the waveform and the two 128x1024 planes.
//
// FFT: a 2048-point real transform is a 1024-point complex FFT of (even, odd) sample pairs plus a
// split step.  One warp owns one frame: each lane holds 32 complex points in registers, does a
// radix-32 pass (5 in-register radix-2 stages), exchanges through a padded per-warp shared-memory
// tile, and does the second radix-32 pass.  The inverse uses the same routine with re/im swapped.
#include "common.cuh"
#include "gansynth_b200.h"
#include <math.h>

namespace {

constexpr int NBINS = 1024;
constexpr int FRAME = 2048;
constexpr int HOP = 512;
constexpr int MEL_TAPS = 6;

__device__ __forceinline__ constexpr int brev5(int k) {
  return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4);
}

// cos / sin of 2*pi*q/32
__device__ __forceinline__ constexpr float cos32(int q) {
  return q == 1 ? 0.98078528040323043f : q == 2 ? 0.92387953251128674f : q == 3 ? 0.83146961230254524f
       : q == 5 ? 0.55557023301960218f : q == 6 ? 0.38268343236508978f : q == 7 ? 0.19509032201612825f
       : q == 9 ? -0.19509032201612825f : q == 10 ? -0.38268343236508978f : q == 11 ? -0.55557023301960218f
       : q == 13 ? -0.83146961230254524f : q == 14 ? -0.92387953251128674f : q == 15 ? -0.98078528040323043f : 0.0f;
}
__device__ __forceinline__ constexpr float sin32(int q) {
  return q == 1 ? 0.19509032201612825f : q == 2 ? 0.38268343236508978f : q == 3 ? 0.55557023301960218f
       : q == 5 ? 0.83146961230254524f : q == 6 ? 0.92387953251128674f : q == 7 ? 0.98078528040323043f
       : q == 9 ? 0.98078528040323043f : q == 10 ? 0.92387953251128674f : q == 11 ? 0.83146961230254524f
       : q == 13 ? 0.55557023301960218f : q == 14 ? 0.38268343236508978f : q == 15 ? 0.19509032201612825f : 0.0f;
}

// In-register 32-point forward DFT (decimation in frequency).  Output index k is left in slot brev5(k).
__device__ __forceinline__ void fft32(float (&re)[32], float (&im)[32]) {
  constexpr float R = 0.70710678118654752f;
#pragma unroll
  for (int lg = 0; lg < 5; ++lg) {
    const int s = 16 >> lg;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if ((i & s) == 0) {
        const int j = i + s;
        const int q = (i & (s - 1)) * (16 / s);
        const float ar = re[i], ai = im[i], br = re[j], bi = im[j];
        re[i] = ar + br;
        im[i] = ai + bi;
        const float dr = ar - br, di = ai - bi;
        if (q == 0) { re[j] = dr; im[j] = di; }
        else if (q == 8) { re[j] = di; im[j] = -dr; }
        else if (q == 4) { re[j] = (dr + di) * R; im[j] = (di - dr) * R; }
        else if (q == 12) { re[j] = (di - dr) * R; im[j] = -(dr + di) * R; }
        else {
          const float c = cos32(q), sn = sin32(q);  // multiply by c - i*sn
          re[j] = dr * c + di * sn;
          im[j] = di * c - dr * sn;
        }
      }
    }
  }
}

// atan2 with the Abramowitz-Stegun 4.4.49 polynomial (|error| <= 2e-8 on [0, 1], below fp32 resolution of the result):
// ~25 instructions instead of the ~60 of atan2f.  atan2(+0, x < 0) = pi, atan2(0, 0) = 0 like the library function.
__device__ __forceinline__ float gs_atan2f(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float a = (mx > 0.0f) ? __fdividef(mn, mx) : 0.0f;
  const float s = a * a;
  float r = 0.0028662257f;
  r = fmaf(r, s, -0.0161657367f);
  r = fmaf(r, s, 0.0429096138f);
  r = fmaf(r, s, -0.0752896400f);
  r = fmaf(r, s, 0.1065626393f);
  r = fmaf(r, s, -0.1420889944f);
  r = fmaf(r, s, 0.1999355085f);
  r = fmaf(r, s, -0.3333314528f);
  r = fmaf(r * s, a, a);
  if (ay > ax) r = 1.57079637050628662f - r;
  if (x < 0.0f) r = 3.14159274101257324f - r;
  return (y < 0.0f) ? -r : r;
}

constexpr int XPITCH = 33;
constexpr int PLANE = 32 * XPITCH;  // 1056 floats
constexpr int WARP_BUF = 2 * PLANE;

// 1024-point forward DFT over one warp.  In: lane holds z[32*n1 + lane] in slot n1.
// Out: lane holds Z[lane + 32*k2] in slot brev5(k2).  tw[k1*32 + lane] = exp(-2*pi*i*k1*lane/1024).
__device__ __forceinline__ void warp_fft1024(float (&re)[32], float (&im)[32], float* buf, const float2* tw, int lane) {
  fft32(re, im);
  float* br = buf;
  float* bi = buf + PLANE;
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) {
    const int r = brev5(k1);
    const float2 w = tw[k1 * 32 + lane];
    const float yr = re[r], yi = im[r];
    br[k1 * XPITCH + lane] = yr * w.x - yi * w.y;
    bi[k1 * XPITCH + lane] = yr * w.y + yi * w.x;
  }
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) {
    re[n2] = br[lane * XPITCH + n2];
    im[n2] = bi[lane * XPITCH + n2];
  }
  __syncwarp();
  fft32(re, im);
}

struct Tables {
  float2* tw1024;  // [32*32]
  float2* tw2048;  // [1024]
  bool ready;
};
Tables g_tables = {nullptr, nullptr, false};

int ensure_tables(cudaStream_t st) {
  if (g_tables.ready) return GS_OK;
  static float2 h1[1024], h2[1024];
  for (int k1 = 0; k1 < 32; ++k1)
    for (int l = 0; l < 32; ++l) {
      double a = -2.0 * M_PI * (double)(k1 * l) / 1024.0;
      h1[k1 * 32 + l] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int k = 0; k < 1024; ++k) {
    double a = -2.0 * M_PI * (double)k / 2048.0;
    h2[k] = make_float2((float)cos(a), (float)sin(a));
  }
  GS_CUDA(cudaMalloc(&g_tables.tw1024, sizeof(h1)));
  GS_CUDA(cudaMalloc(&g_tables.tw2048, sizeof(h2)));
  GS_CUDA(cudaMemcpyAsync(g_tables.tw1024, h1, sizeof(h1), cudaMemcpyHostToDevice, st));
  GS_CUDA(cudaMemcpyAsync(g_tables.tw2048, h2, sizeof(h2), cudaMemcpyHostToDevice, st));
  GS_CUDA(cudaStreamSynchronize(st));
  g_tables.ready = true;
  return GS_OK;
}

// ------------------------------------------------------------------------------------------------
// Forward: waveform -> (log-mel magnitude, mel instantaneous frequency).
// One warp per (clip, chunk of L consecutive frames); a chunk that does not start the clip first
// recomputes frame t0-1 to get the previous mel phase.
constexpr int FWD_WARPS = 16;   // 16 x 12.25 KB of per-warp tiles + 16 KB of tables = 212 KB: one CTA of 16 warps per SM
constexpr int FWD_WARP_FLOATS = WARP_BUF + NBINS;  // FFT tile (reused for mag/phase) + previous mel phase
constexpr int FWD_SMEM = (2 * 1024 * 2 + FWD_WARPS * FWD_WARP_FLOATS) * 4;

__global__ void __launch_bounds__(FWD_WARPS * 32)
spectrogram_fwd_kernel(const float* __restrict__ wave, int wave_len, int T, int L, const float* __restrict__ hann,
                       const float2* __restrict__ tw1024g, const float2* __restrict__ tw2048g,
                       const int* __restrict__ mel_k0, const float* __restrict__ mel_w, float* __restrict__ logmel,
                       float* __restrict__ inst, int n_items) {
  extern __shared__ __align__(16) float sm[];
  float2* tw1024 = reinterpret_cast<float2*>(sm);
  float2* tw2048 = tw1024 + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* buf = sm + 4096 + warp * FWD_WARP_FLOATS;
  float* prev = buf + WARP_BUF;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    tw1024[i] = tw1024g[i];
    tw2048[i] = tw2048g[i];
  }
  __syncthreads();
  const int item = blockIdx.x * FWD_WARPS + warp;
  if (item >= n_items) return;
  const int chunks = T / L;
  const int b = item / chunks;
  const int t0 = (item % chunks) * L;
  const int padf = HOP * (T - 1) + FRAME - wave_len;
  const float* wv = wave + (size_t)b * wave_len;
  const float PI_F = 3.14159274101257324f;
  const float TWO_PI_F = 6.28318548202514648f;
  float* zr = buf;
  float* zi = buf + PLANE;

  for (int t = (t0 > 0 ? t0 - 1 : 0); t < t0 + L; ++t) {
    float re[32], im[32];
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) {
      const int nn = 64 * n1 + 2 * lane;
      const int s = HOP * t + nn - padf;
      const float2 hw = __ldg(reinterpret_cast<const float2*>(hann + nn));
      float v0 = (s >= 0 && s < wave_len) ? __ldg(wv + s) : 0.0f;
      float v1 = (s + 1 >= 0 && s + 1 < wave_len) ? __ldg(wv + s + 1) : 0.0f;
      re[n1] = v0 * hw.x;
      im[n1] = v1 * hw.y;
    }
    warp_fft1024(re, im, buf, tw1024, lane);
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) {
      zr[lane + 32 * k2] = re[brev5(k2)];
      zi[lane + 32 * k2] = im[brev5(k2)];
    }
    __syncwarp();
    // split step: X[k] = E + W^k O, X[1024-k] = conj(E - W^k O); magnitude -> zr, phase -> zi, stored at slot (k & 1023)
    for (int m = 0; m <= 16; ++m) {
      const int k = lane + 32 * m;
      if (k <= 512) {
        const int kc = (1024 - k) & 1023;
        const float ar = zr[k], ai = zi[k], cr = zr[kc], ci = zi[kc];
        const float er = 0.5f * (ar + cr), ei = 0.5f * (ai - ci);
        const float orr = 0.5f * (ai + ci), oi = -0.5f * (ar - cr);
        const float2 w = tw2048[k];
        const float tr = orr * w.x - oi * w.y, ti = orr * w.y + oi * w.x;
        const float xr = er + tr, xi = ei + ti;        // X[k]
        const float yr = er - tr, yi = -(ei - ti);     // X[1024-k]
        if (k >= 1) {
          zr[k] = sqrtf(fmaf(xr, xr, xi * xi));
          zi[k] = gs_atan2f(xi + 0.0f, xr + 0.0f);
        }
        if (k != 512) {
          zr[kc] = sqrtf(fmaf(yr, yr, yi * yi));
          zi[kc] = gs_atan2f(yi + 0.0f, yr + 0.0f);
        }
      }
    }
    __syncwarp();
    const bool emit = (t >= t0);
    float* lm_out = logmel + ((size_t)b * T + t) * NBINS;
    float* if_out = inst + ((size_t)b * T + t) * NBINS;
#pragma unroll 4
    for (int m = 0; m < 32; ++m) {
      const int j = lane + 32 * m;
      const int k0 = __ldg(mel_k0 + j);
      float mm = 0.0f, pp = 0.0f;
#pragma unroll
      for (int i = 0; i < MEL_TAPS; ++i) {
        const float wgt = __ldg(mel_w + i * NBINS + j);
        const int slot = (k0 + i + 1) & 1023;
        mm = fmaf(zr[slot], wgt, mm);
        pp = fmaf(zi[slot], wgt, pp);
      }
      if (emit) {
        lm_out[j] = (__logf(mm + 1.0e-6f) + 3.76f) * (1.0f / 10.05f);
        float v;
        if (t == 0) {
          v = pp * 0.318309886183790672f;
        } else {
          const float d = pp - prev[j];
          // floor-mod of d + pi by 2 pi (|d| < 2 pi: one correction step each way covers the rounding of the quotient)
          const float tt = d + PI_F;
          float md = fmaf(-floorf(tt * 0.159154943091895336f), TWO_PI_F, tt);
          if (md < 0.0f) md += TWO_PI_F;
          if (md >= TWO_PI_F) md -= TWO_PI_F;
          md -= PI_F;
          if (md == -PI_F && d > 0.0f) md = PI_F;
          v = md * 0.318309886183790672f;
        }
        if_out[j] = v;
      }
      prev[j] = pp;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Inverse: (log-mel magnitude, mel IF) -> waveform.  One CTA per clip, 8 warps, frames in groups of 8:
// running phase cumsum in registers, banded mel->linear product, one inverse FFT per warp, overlap-add
// through a carry buffer so every output sample is written exactly once.
constexpr int INV_F = 8;
constexpr int INV_THREADS = INV_F * 32;
constexpr int CARRY = FRAME - HOP;  // 1536
constexpr int XLEN = 1028;          // X planes hold k = 0..1024
constexpr int INV_SMEM_FLOATS = 4096 /*tables*/ + 2 * INV_F * NBINS /*mel mag/phase*/ + 2 * INV_F * XLEN /*X*/ +
                                INV_F * WARP_BUF /*fft tiles + frames*/ + CARRY;
constexpr int INV_SMEM = INV_SMEM_FLOATS * 4;

__global__ void __launch_bounds__(INV_THREADS, 1)
waveform_fwd_kernel(const float* __restrict__ logmel, const float* __restrict__ inst, int T, int wave_len,
                    const float* __restrict__ synwin, const float2* __restrict__ tw1024g,
                    const float2* __restrict__ tw2048g, const int* __restrict__ pb_j0, const int* __restrict__ pb_cnt,
                    const float* __restrict__ pb_w, int band, float* __restrict__ wave) {
  extern __shared__ __align__(16) float sm[];
  float2* tw1024 = reinterpret_cast<float2*>(sm);
  float2* tw2048 = tw1024 + 1024;
  float* mmag = sm + 4096;                    // [F][1024]
  float* mph = mmag + INV_F * NBINS;          // [F][1024]
  float* xr = mph + INV_F * NBINS;            // [F][XLEN]
  float* xi = xr + INV_F * XLEN;              // [F][XLEN]
  float* tiles = xi + INV_F * XLEN;           // [F][WARP_BUF]
  float* carry = tiles + INV_F * WARP_BUF;    // [1536]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  for (int i = tid; i < 1024; i += INV_THREADS) {
    tw1024[i] = tw1024g[i];
    tw2048[i] = tw2048g[i];
  }
  for (int i = tid; i < CARRY; i += INV_THREADS) carry[i] = 0.0f;
  const float PI_F = 3.14159274101257324f;
  const int padf = HOP * (T - 1) + FRAME - wave_len;
  float run[4] = {0.f, 0.f, 0.f, 0.f};
  float* out = wave + (size_t)b * wave_len;
  __syncthreads();

  for (int g0 = 0; g0 < T; g0 += INV_F) {
    // (a) un-normalise, exp, running phase
    {
      float lmv[INV_F][4], ifv[INV_F][4];
#pragma unroll
      for (int f = 0; f < INV_F; ++f)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const size_t o = ((size_t)b * T + g0 + f) * NBINS + tid + INV_THREADS * q;
          const bool ok = (g0 + f) < T;
          lmv[f][q] = ok ? __ldg(logmel + o) : 0.0f;
          ifv[f][q] = ok ? __ldg(inst + o) : 0.0f;
        }
#pragma unroll
      for (int f = 0; f < INV_F; ++f)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = tid + INV_THREADS * q;
          run[q] = __fadd_rn(run[q], __fmul_rn(ifv[f][q], PI_F));
          mph[f * NBINS + j] = run[q];
          mmag[f * NBINS + j] = expf(__fadd_rn(__fmul_rn(lmv[f][q], 10.05f), -3.76f));
        }
    }
    __syncthreads();
    // (b) banded mel -> linear for magnitude and phase, then polar -> rectangular
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
      const int d = tid + INV_THREADS * q;
      const int j0 = __ldg(pb_j0 + d);
      int cnt = __ldg(pb_cnt + d);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt = max(cnt, __shfl_xor_sync(0xffffffffu, cnt, o));
      float am[INV_F], ap[INV_F];
#pragma unroll
      for (int f = 0; f < INV_F; ++f) { am[f] = 0.0f; ap[f] = 0.0f; }
      // the band weights come from L2: fetch eight at a time so their latencies overlap (rows >= this lane's own
      // count hold zeros, the warp-wide maximum `cnt` only bounds the loop)
      for (int i0 = 0; i0 < cnt; i0 += 8) {
        float c8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) c8[u] = (i0 + u < band) ? __ldg(pb_w + (size_t)(i0 + u) * NBINS + d) : 0.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = min(j0 + i0 + u, NBINS - 1);
          const float c = c8[u];
#pragma unroll
          for (int f = 0; f < INV_F; ++f) {
            am[f] = fmaf(mmag[f * NBINS + j], c, am[f]);
            ap[f] = fmaf(mph[f * NBINS + j], c, ap[f]);
          }
        }
      }
#pragma unroll
      for (int f = 0; f < INV_F; ++f) {
        float sn, cs;
        sincosf(ap[f], &sn, &cs);
        xr[f * XLEN + d + 1] = am[f] * cs;
        xi[f * XLEN + d + 1] = am[f] * sn;
      }
    }
    if (tid < INV_F) { xr[tid * XLEN] = 0.0f; xi[tid * XLEN] = 0.0f; }
    __syncthreads();
    // (c) one inverse real FFT per warp
    {
      const int f = warp;
      const float* fxr = xr + f * XLEN;
      const float* fxi = xi + f * XLEN;
      float* tile = tiles + f * WARP_BUF;
      float re[32], im[32];
#pragma unroll
      for (int n1 = 0; n1 < 32; ++n1) {
        const int k = 32 * n1 + lane;
        const int kc = 1024 - k;
        // X[0] = 0 (DC dropped), X[1024] is taken as real (irfft ignores its imaginary part)
        const float ar = fxr[k], ai = fxi[k];
        const float cr = fxr[kc], ci = (kc == 1024) ? 0.0f : fxi[kc];
        const float ai0 = (k == 0) ? 0.0f : ai;
        // E = (X[k] + conj(X[N-k]))/2 ; D = (X[k] - conj(X[N-k]))/2 ; O = conj(W^k) D ; Z = E + iO
        const float er = 0.5f * (ar + cr), ei = 0.5f * (ai0 - ci);
        const float dr = 0.5f * (ar - cr), di = 0.5f * (ai0 + ci);
        float2 w = tw2048[k & 1023];  // (cos, -sin); k <= 1023 here
        // conj(W^k) = (w.x, -w.y)
        const float orr = dr * w.x + di * w.y, oi = di * w.x - dr * w.y;
        re[n1] = er - oi;
        im[n1] = ei + orr;
      }
      warp_fft1024(im, re, tile, tw1024, lane);  // swapped arguments = inverse transform
      // lane holds z[m], m = lane + 32*k2 : x[2m] = Re/1024, x[2m+1] = Im/1024
      const float sc = 1.0f / 1024.0f;
#pragma unroll
      for (int k2 = 0; k2 < 32; ++k2) {
        const int m = lane + 32 * k2;
        const float2 sw = __ldg(reinterpret_cast<const float2*>(synwin + 2 * m));
        float2 v;
        v.x = (re[brev5(k2)] * sc) * sw.x;
        v.y = (im[brev5(k2)] * sc) * sw.y;
        *reinterpret_cast<float2*>(tile + 2 * m) = v;
      }
    }
    __syncthreads();
    // (d) overlap-add: finished samples [HOP*g0, HOP*(g0+F)), ascending frame order like the reference
    {
      const int nf = (T - g0 < INV_F) ? T - g0 : INV_F;
      const int span = HOP * nf;
      for (int s = tid; s < span + CARRY; s += INV_THREADS) {
        float acc = (s < CARRY) ? carry[s] : 0.0f;
        const int fhi = s / HOP;
        int flo = fhi - 3;
        if (flo < 0) flo = 0;
        for (int f = flo; f <= fhi && f < nf; ++f) acc += tiles[f * WARP_BUF + (s - HOP * f)];
        if (s < span) {
          const int ng = HOP * g0 + s - padf;
          if (ng >= 0 && ng < wave_len) out[ng] = acc;
        } else {
          // becomes the next carry; stash in registers-free way: write after the barrier below
          xr[s - span] = acc;
        }
      }
      __syncthreads();
      for (int s = tid; s < CARRY; s += INV_THREADS) carry[s] = xr[s];
      __syncthreads();
      if (g0 + INV_F >= T) {
        for (int s = tid; s < CARRY; s += INV_THREADS) {
          const int ng = HOP * T + s - padf;
          if (ng >= 0 && ng < wave_len) out[ng] = carry[s];
        }
      }
    }
  }
}

}  // namespace

extern "C" int gs_spectrogram_fwd(const float* wave, const float* hann, const int* mel_k0, const float* mel_w,
                                  float* logmel, float* inst, int batch, int wave_len, int time_steps,
                                  int frames_per_chunk, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GS_CHECK_ARG(batch >= 0 && wave_len > 0 && time_steps > 0, "spectrogram_fwd: bad shape");
  GS_CHECK_ARG(HOP * (time_steps - 1) + FRAME >= wave_len,
               "spectrogram_fwd: waveform_length %d exceeds the %d samples %d frames cover", wave_len,
               HOP * (time_steps - 1) + FRAME, time_steps);
  GS_CHECK_ARG(frames_per_chunk > 0 && time_steps % frames_per_chunk == 0,
               "spectrogram_fwd: frames_per_chunk %d must divide time_steps %d", frames_per_chunk, time_steps);
  if (batch == 0) return GS_OK;
  int rc = ensure_tables(st);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    GS_CUDA(cudaFuncSetAttribute(spectrogram_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    attr = true;
  }
  int n_items = batch * (time_steps / frames_per_chunk);
  int blocks = gs_cdiv(n_items, FWD_WARPS);
  spectrogram_fwd_kernel<<<blocks, FWD_WARPS * 32, FWD_SMEM, st>>>(wave, wave_len, time_steps, frames_per_chunk, hann,
                                                                  g_tables.tw1024, g_tables.tw2048, mel_k0, mel_w,
                                                                  logmel, inst, n_items);
  GS_CHECK_LAUNCH("spectrogram_fwd");
  return GS_OK;
}

extern "C" int gs_waveform_fwd(const float* logmel, const float* inst, const float* synth_window, const int* pb_j0,
                               const int* pb_cnt, const float* pb_w, int band, float* wave, int batch, int wave_len,
                               int time_steps, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GS_CHECK_ARG(batch >= 0 && wave_len > 0 && time_steps > 0 && band > 0, "waveform_fwd: bad shape");
  GS_CHECK_ARG(HOP * (time_steps - 1) + FRAME >= wave_len,
               "waveform_fwd: waveform_length %d exceeds the %d samples %d frames cover", wave_len,
               HOP * (time_steps - 1) + FRAME, time_steps);
  if (batch == 0) return GS_OK;
  int rc = ensure_tables(st);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    GS_CUDA(cudaFuncSetAttribute(waveform_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, INV_SMEM));
    attr = true;
  }
  waveform_fwd_kernel<<<batch, INV_THREADS, INV_SMEM, st>>>(logmel, inst, time_steps, wave_len, synth_window,
                                                           g_tables.tw1024, g_tables.tw2048, pb_j0, pb_cnt, pb_w, band,
                                                           wave);
  GS_CHECK_LAUNCH("waveform_fwd");
  return GS_OK;
}
