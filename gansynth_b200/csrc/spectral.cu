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

// Twiddle tables tw1024[32 * 32] and tw2048[1024] (float2 each) live at the head of the caller's workspace
// (gs_context, common.cuh).  They are computed ON THE DEVICE by a kernel enqueued on the stream of the context's first
// spectral call -- in double precision, rounded to float -- so no entry point allocates, copies from the host or
// synchronises.
__global__ void twiddle_init_kernel(float2* __restrict__ tw1024, float2* __restrict__ tw2048) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 1024) {
    double sn, cs;
    sincospi(-2.0 * (double)((i >> 5) * (i & 31)) / 1024.0, &sn, &cs);
    tw1024[i] = make_float2((float)cs, (float)sn);
    sincospi(-2.0 * (double)i / 2048.0, &sn, &cs);
    tw2048[i] = make_float2((float)cs, (float)sn);
  }
}

struct Tables {
  float2* tw1024;  // [32*32]
  float2* tw2048;  // [1024]
};

int ensure_tables(gs_context* ctx, cudaStream_t st, Tables* t) {
  t->tw1024 = reinterpret_cast<float2*>(ctx->ws + ctx->tables_off);
  t->tw2048 = t->tw1024 + 1024;
  if (ctx->tables_ready) return GS_OK;
  twiddle_init_kernel<<<4, 256, 0, st>>>(t->tw1024, t->tw2048);
  GS_CHECK_LAUNCH("twiddle_init");
  ctx->tables_ready = true;
  return GS_OK;
}

// ------------------------------------------------------------------------------------------------
// Forward: waveform -> (log-mel magnitude, mel instantaneous frequency).
//
// One CTA owns a RUN of consecutive frames of one clip and walks it in lock-step rounds of FWD_WARPS frames,
// one frame per warp: neighbouring warps read overlapping samples at the same time (L1 reuse of the 75 %
// overlap), and the mel phase of frame t-1 that the instantaneous frequency of frame t needs is the
// neighbouring warp's result of the same round (shared memory), never a recomputed transform.  Only the first
// frame of a run (t0 > 0) has no neighbour: it is written as raw phase, the last frame of the run before it
// leaves its phase in `scratch`, and spectrogram_if_fixup_kernel finishes those rows.
//
// Per warp: an 8 KB tile = XOR-swizzled FFT exchange planes, then the 1024 (magnitude, phase) pairs of the
// frame (slot d = FFT bin d + 1: the DC bin is dropped, the Nyquist bin is slot 1023), then the frame's 1024
// mel phases for the neighbour.  All constant tables (mel taps, twiddles, Hann) sit in shared memory.
constexpr int FWD_WARPS = 16;
constexpr int TILE = 2048;                       // floats per warp
constexpr int FWD_TABLE_FLOATS = NBINS /*mel k0*/ + MEL_TAPS * NBINS /*mel w*/ + 2048 /*tw1024*/ + 1040 /*tw2048, k<=512*/ +
                                 FRAME /*hann*/ + 2 * NBINS /*carry*/ + 32 /*taps per 32-bin block*/;
constexpr int FWD_SMEM = (FWD_TABLE_FLOATS + FWD_WARPS * TILE) * 4;

__device__ __forceinline__ float gs_sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// wrapped phase difference in units of pi: ((d + pi) floor-mod 2 pi) - pi, with numpy.unwrap's +pi convention
// (spectral_ops.py:20-31 followed by diff, :34-42).  |d| < 2 pi * 6: one correction step each way covers the rounding
// of the quotient.
__device__ __forceinline__ float gs_wrap_diff(float d) {
  const float PI_F = 3.14159274101257324f;
  const float TWO_PI_F = 6.28318548202514648f;
  const float tt = d + PI_F;
  float md = fmaf(-floorf(tt * 0.159154943091895336f), TWO_PI_F, tt);
  if (md < 0.0f) md += TWO_PI_F;
  if (md >= TWO_PI_F) md -= TWO_PI_F;
  md -= PI_F;
  if (md == -PI_F && d > 0.0f) md = PI_F;
  return md * 0.318309886183790672f;
}

// 1024-point forward DFT over one warp, exchange through one XOR-swizzled 32x32 tile of complex values (64-bit
// accesses, conflict-free both ways).  In: lane holds z[32*n1 + lane] in slot n1.  Out: lane holds Z[lane + 32*k2] in
// slot brev5(k2).
__device__ __forceinline__ void warp_fft1024_sw(float (&re)[32], float (&im)[32], float2* buf, const float2* tw, int lane) {
  fft32(re, im);
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) {
    const int r = brev5(k1);
    const float2 w = tw[k1 * 32 + lane];
    const float yr = re[r], yi = im[r];
    buf[k1 * 32 + (lane ^ k1)] = make_float2(yr * w.x - yi * w.y, yr * w.y + yi * w.x);
  }
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) {
    const float2 v = buf[lane * 32 + (n2 ^ lane)];
    re[n2] = v.x;
    im[n2] = v.y;
  }
  __syncwarp();
  fft32(re, im);
}

__global__ void __launch_bounds__(FWD_WARPS * 32, 1)
spectrogram_fwd_kernel(const float* __restrict__ wave, int wave_len, int T, int RL, int runs_per_clip,
                       const float* __restrict__ hann_g, const float2* __restrict__ tw1024g,
                       const float2* __restrict__ tw2048g, const int* __restrict__ mel_k0,
                       const float* __restrict__ mel_w, float* __restrict__ logmel, float* __restrict__ inst,
                       float* __restrict__ scratch, int total_runs) {
  extern __shared__ __align__(16) float sm[];
  int* melk = reinterpret_cast<int*>(sm);                           // [1024]
  float* melw = sm + NBINS;                                         // [6][1024]
  float2* tw1024 = reinterpret_cast<float2*>(melw + MEL_TAPS * NBINS);  // [1024]
  float2* tw2048 = tw1024 + 1024;                                   // [513] (+pad)
  float2* hann2 = reinterpret_cast<float2*>(sm + NBINS + MEL_TAPS * NBINS + 2048 + 1040);  // [1024]
  float* carry = reinterpret_cast<float*>(hann2 + 1024);            // [2][1024]
  int* ntaps = reinterpret_cast<int*>(carry + 2 * NBINS);           // [32]
  float* tiles = carry + 2 * NBINS + 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* buf = tiles + warp * TILE;
  for (int i = tid; i < 1024; i += FWD_WARPS * 32) {
    // taps of mel bin i start at data row k0; keep all MEL_TAPS reads inside [0, 1024) by sliding the window down
    const int k0 = __ldg(mel_k0 + i);
    const int k0c = min(max(k0, 0), NBINS - MEL_TAPS);
    const int sh = k0 - k0c;
    melk[i] = k0c;
    int used = 0;
#pragma unroll
    for (int u = 0; u < MEL_TAPS; ++u) {
      const int src = u - sh;
      const float wv_ = (src >= 0 && src < MEL_TAPS) ? __ldg(mel_w + src * NBINS + i) : 0.0f;
      melw[u * NBINS + i] = wv_;
      if (wv_ != 0.0f) used = u + 1;
    }
    // the filters are 1 row wide at the bottom of the mel axis and 6 at the top: taps beyond the widest filter of a
    // 32-bin block (one warp-wide step of the projection) are skipped
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) used = max(used, __shfl_xor_sync(0xffffffffu, used, o));
    if (lane == 0) ntaps[i >> 5] = used;
    tw1024[i] = tw1024g[i];
    if (i <= 512) tw2048[i] = tw2048g[i];
    hann2[i] = __ldg(reinterpret_cast<const float2*>(hann_g) + i);
  }
  __syncthreads();

  const int padf = HOP * (T - 1) + FRAME - wave_len;
  const bool vec_ok = (((wave_len | padf) & 1) == 0) && ((reinterpret_cast<uintptr_t>(wave) & 7) == 0);
  float2* zc = reinterpret_cast<float2*>(buf);

  // persistent CTAs: the 48 KB of tables are loaded once per SM, not once per run
  for (int run_id = blockIdx.x; run_id < total_runs; run_id += gridDim.x) {
  const int b = run_id / runs_per_clip;
  const int run = run_id % runs_per_clip;
  const int t0 = run * RL;
  const int t1 = min(T, t0 + RL);
  const float* wv = wave + (size_t)b * wave_len;
  int round = 0;
  for (int tb = t0; tb < t1; tb += FWD_WARPS, ++round) {
    const int t = tb + warp;
    const bool active = t < t1;
    float pp[32];
    if (active) {
      float re[32], im[32];
      const int base = HOP * t - padf;  // sample index of frame element 0
      if (vec_ok && base >= 0 && base + FRAME <= wave_len) {
        const float2* p = reinterpret_cast<const float2*>(wv + base) + lane;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
          const float2 v = __ldg(p + 32 * n1);
          const float2 hw = hann2[32 * n1 + lane];
          re[n1] = v.x * hw.x;
          im[n1] = v.y * hw.y;
        }
      } else {
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
          const int s = base + 64 * n1 + 2 * lane;
          const float2 hw = hann2[32 * n1 + lane];
          const float v0 = (s >= 0 && s < wave_len) ? __ldg(wv + s) : 0.0f;
          const float v1 = (s + 1 >= 0 && s + 1 < wave_len) ? __ldg(wv + s + 1) : 0.0f;
          re[n1] = v0 * hw.x;
          im[n1] = v1 * hw.y;
        }
      }
      warp_fft1024_sw(re, im, zc, tw1024, lane);
#pragma unroll
      for (int k2 = 0; k2 < 32; ++k2) zc[lane + 32 * k2] = make_float2(re[brev5(k2)], im[brev5(k2)]);
      __syncwarp();
      // split step: X[k] = E + W^k O, X[1024-k] = conj(E - W^k O) from Z[k], Z[1024-k]; all operands are read
      // before any (magnitude, phase) pair is written because X[k] lands in slot k-1
      float2 za[17], zb[17];
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const int k = lane + 32 * m;
        za[m] = zc[k];
        zb[m] = zc[(1024 - k) & 1023];
      }
      za[16] = zc[512];
      zb[16] = za[16];
      __syncwarp();
#pragma unroll
      for (int m = 0; m < 17; ++m) {
        const int k = (m < 16) ? lane + 32 * m : 512;
        const float ar = za[m].x, ai = za[m].y, cr = zb[m].x, ci = zb[m].y;
        const float er = 0.5f * (ar + cr), ei = 0.5f * (ai - ci);
        const float orr = 0.5f * (ai + ci), oi = -0.5f * (ar - cr);
        const float2 w = tw2048[k];
        const float tr = orr * w.x - oi * w.y, ti = orr * w.y + oi * w.x;
        const float xr = er + tr, xi = ei + ti;        // X[k]
        const float yr = er - tr, yi = ti - ei;        // X[1024-k]
        if (m < 16) {
          if (k >= 1) zc[k - 1] = make_float2(gs_sqrt_approx(fmaf(xr, xr, xi * xi)), gs_atan2f(xi + 0.0f, xr + 0.0f));
          zc[1023 - k] = make_float2(gs_sqrt_approx(fmaf(yr, yr, yi * yi)), gs_atan2f(yi + 0.0f, yr + 0.0f));
        } else if (lane == 0) {
          zc[511] = make_float2(gs_sqrt_approx(fmaf(xr, xr, xi * xi)), gs_atan2f(xi + 0.0f, xr + 0.0f));
        }
      }
      __syncwarp();
      // sparse mel projection of magnitude and phase (spectral_ops.py:76-85), log-magnitude out
      float* lm_out = logmel + ((size_t)b * T + t) * NBINS;
#pragma unroll
      for (int m = 0; m < 32; ++m) {
        const int j = lane + 32 * m;
        const float2* z = zc + melk[j];
        const int nt = ntaps[m];
        float mm = 0.0f, ph = 0.0f;
#pragma unroll
        for (int i = 0; i < MEL_TAPS; ++i) {
          if (i >= nt) break;
          const float wgt = melw[i * NBINS + j];
          const float2 v = z[i];
          mm = fmaf(v.x, wgt, mm);
          ph = fmaf(v.y, wgt, ph);
        }
        pp[m] = ph;
        lm_out[j] = (__logf(mm + 1.0e-6f) + 3.76f) * (1.0f / 10.05f);
      }
      __syncwarp();
      // publish this frame's mel phase for the next frame's warp (the last warp feeds warp 0 of the next round)
      float* pub = (warp == FWD_WARPS - 1) ? carry + (round & 1) * NBINS : buf;
#pragma unroll
      for (int m = 0; m < 32; ++m) pub[lane + 32 * m] = pp[m];
    }
    __syncthreads();
    if (active) {
      float* if_out = inst + ((size_t)b * T + t) * NBINS;
      if (t == 0) {
#pragma unroll
        for (int m = 0; m < 32; ++m) if_out[lane + 32 * m] = pp[m] * 0.318309886183790672f;
      } else if (t == t0) {
        // first frame of a later run: raw phase, finished by spectrogram_if_fixup_kernel
#pragma unroll
        for (int m = 0; m < 32; ++m) if_out[lane + 32 * m] = pp[m];
      } else {
        const float* prev = (warp == 0) ? carry + ((round + 1) & 1) * NBINS : buf - TILE;
#pragma unroll
        for (int m = 0; m < 32; ++m) if_out[lane + 32 * m] = gs_wrap_diff(pp[m] - prev[lane + 32 * m]);
      }
      if (t == t1 - 1 && t1 < T) {
        float* sc = scratch + ((size_t)b * runs_per_clip + run) * NBINS;
#pragma unroll
        for (int m = 0; m < 32; ++m) sc[lane + 32 * m] = pp[m];
      }
    }
    __syncthreads();
  }
  }
}

// Rows t0 = run * RL (run >= 1) hold raw mel phase; scratch[b][run-1] holds the mel phase of frame t0 - 1.
__global__ void __launch_bounds__(256)
spectrogram_if_fixup_kernel(float* __restrict__ inst, const float* __restrict__ scratch, int T, int RL,
                            int runs_per_clip, int rows) {
  const int row = blockIdx.x;
  if (row >= rows) return;
  const int b = row / (runs_per_clip - 1);
  const int run = row % (runs_per_clip - 1) + 1;
  float4* cur = reinterpret_cast<float4*>(inst + ((size_t)b * T + (size_t)run * RL) * NBINS) + threadIdx.x;
  const float4 p = __ldg(reinterpret_cast<const float4*>(scratch + ((size_t)b * runs_per_clip + run - 1) * NBINS) + threadIdx.x);
  float4 c = *cur;
  c.x = gs_wrap_diff(c.x - p.x);
  c.y = gs_wrap_diff(c.y - p.y);
  c.z = gs_wrap_diff(c.z - p.z);
  c.w = gs_wrap_diff(c.w - p.w);
  *cur = c;
}

// ------------------------------------------------------------------------------------------------
// Inverse: (log-mel magnitude, mel IF) -> waveform.  One CTA of 16 warps per clip, frames in groups of 8, three
// stages per group separated by CTA barriers:
//   (b) all warps: banded mel->linear product for magnitude and phase + polar->rectangular.  The 8 frames of a
//       mel bin sit in ONE padded shared-memory row (8 magnitudes, 8 phases), so a band tap costs 4 LDS.128 for
//       16 FMAs; the 32 column blocks are handed out heaviest-first through a shared counter because the band is
//       2 rows wide at low frequencies and 46 around bin 330;
//   (c) warps 0-7: one inverse real FFT each (spectrum and frame share the warp's tile), synthesis window;
//       warps 8-15 meanwhile do stage (a) of the NEXT group: un-normalise, exp, running phase cumsum in registers;
//   (d) all warps: overlap-add through a ping-pong carry so every output sample is written exactly once.
constexpr int INV_F = 8;
constexpr int INV_WARPS = 16;
constexpr int INV_THREADS = INV_WARPS * 32;
constexpr int CARRY = FRAME - HOP;  // 1536
constexpr int XLEN = 1028;          // X planes hold k = 0..1024
constexpr int MROW = 20;            // floats per mel row: 8 mag + 8 phase + 4 pad (8 consecutive rows hit 8 distinct bank quads)
constexpr int INV_SMEM_FLOATS = 4096 /*tables*/ + NBINS * MROW + INV_F * WARP_BUF /*X, fft tiles, frames*/ + 2 * CARRY + 64;
constexpr int INV_SMEM = INV_SMEM_FLOATS * 4;

// Phase prefix of every later segment: prefix[b][seg][j] = sum_{t < seg * seg_frames} inst[b][t][j] * pi, accumulated
// in frame order with the same two roundings per step as the in-kernel cumsum (tf.cumsum, spectral_ops.py:112).
__global__ void __launch_bounds__(256)
waveform_phase_prefix_kernel(const float* __restrict__ inst, int T, int seg_frames, int segs, float* __restrict__ prefix) {
  const int b = blockIdx.x >> 2;
  const int j = (blockIdx.x & 3) * 256 + threadIdx.x;
  const float PI_F = 3.14159274101257324f;
  const float* col = inst + (size_t)b * T * NBINS + j;
  float run = 0.0f;
  const int t_last = min(T, (segs - 1) * seg_frames);
  for (int t0 = 0; t0 < t_last; t0 += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (t0 + u < t_last) ? __ldg(col + (size_t)(t0 + u) * NBINS) : 0.0f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (t0 + u < t_last) {
        run = __fadd_rn(run, __fmul_rn(v[u], PI_F));
        if ((t0 + u + 1) % seg_frames == 0) prefix[((size_t)b * segs + (t0 + u + 1) / seg_frames) * NBINS + j] = run;
      }
    }
  }
}

__global__ void __launch_bounds__(INV_THREADS, 1)
waveform_fwd_kernel(const float* __restrict__ logmel, const float* __restrict__ inst, int T, int wave_len,
                    const float* __restrict__ synwin, const float2* __restrict__ tw1024g,
                    const float2* __restrict__ tw2048g, const int* __restrict__ pb_j0, const int* __restrict__ pb_cnt,
                    const float* __restrict__ pb_w, int band, float* __restrict__ wave, int seg_frames, int segs,
                    const float* __restrict__ prefix) {
  extern __shared__ __align__(16) float sm[];
  float2* tw1024 = reinterpret_cast<float2*>(sm);
  float2* tw2048 = tw1024 + 1024;
  float* melmp = sm + 4096;                    // [1024][MROW]
  float* tiles = melmp + NBINS * MROW;         // [F][WARP_BUF]
  float* carry = tiles + INV_F * WARP_BUF;     // [2][1536]
  int* order = reinterpret_cast<int*>(carry + 2 * CARRY);  // [32] column blocks, heaviest first
  int* ctr = order + 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // segs > 1 (small batches): the clip is cut into segments of seg_frames frames, one CTA each.  A segment starts
  // from the phase prefix of the frames before it (waveform_phase_prefix_kernel) and exchanges the 1536-sample
  // overlap with its neighbours through atomicAdd on the pre-zeroed output (two addends per sample: order-free).
  const int b = blockIdx.x / segs;
  const int seg = blockIdx.x % segs;
  const int t_begin = seg * seg_frames;
  const int t_end = min(T, t_begin + seg_frames);
  for (int i = tid; i < 1024; i += INV_THREADS) {
    tw1024[i] = tw1024g[i];
    tw2048[i] = tw2048g[i];
  }
  for (int i = tid; i < 2 * CARRY; i += INV_THREADS) carry[i] = 0.0f;
  if (warp == 0) {
    int wmax = 0;
    for (int i = 0; i < 32; ++i) wmax = max(wmax, __ldg(pb_cnt + lane * 32 + i));
    int rank = 0;
    for (int m = 0; m < 32; ++m) {
      const int o = __shfl_sync(0xffffffffu, wmax, m);
      rank += (o > wmax || (o == wmax && m < lane)) ? 1 : 0;
    }
    order[rank] = lane;
    if (lane == 0) *ctr = 0;
  }
  const float PI_F = 3.14159274101257324f;
  const int padf = HOP * (T - 1) + FRAME - wave_len;
  float run[4] = {0.f, 0.f, 0.f, 0.f};
  if (seg > 0 && tid >= 256) {
#pragma unroll
    for (int q = 0; q < 4; ++q) run[q] = __ldg(prefix + ((size_t)b * segs + seg) * NBINS + (tid - 256) + 256 * q);
  }
  float* out = wave + (size_t)b * wave_len;

  // (a) un-normalise, exp, running phase (threads 256..511; column j = u + 256 q)
  auto stage_a = [&](int g0) {
    const int u = tid - 256;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = u + 256 * q;
      float lmv[INV_F], ifv[INV_F];
#pragma unroll
      for (int f = 0; f < INV_F; ++f) {
        const size_t o = ((size_t)b * T + g0 + f) * NBINS + j;
        const bool ok = (g0 + f) < t_end;
        lmv[f] = ok ? __ldg(logmel + o) : 0.0f;
        ifv[f] = ok ? __ldg(inst + o) : 0.0f;
      }
      float mg[INV_F], ph[INV_F];
#pragma unroll
      for (int f = 0; f < INV_F; ++f) {
        run[q] = __fadd_rn(run[q], __fmul_rn(ifv[f], PI_F));
        ph[f] = run[q];
        mg[f] = expf(__fadd_rn(__fmul_rn(lmv[f], 10.05f), -3.76f));
      }
      float4* row = reinterpret_cast<float4*>(melmp + j * MROW);
      row[0] = make_float4(mg[0], mg[1], mg[2], mg[3]);
      row[1] = make_float4(mg[4], mg[5], mg[6], mg[7]);
      row[2] = make_float4(ph[0], ph[1], ph[2], ph[3]);
      row[3] = make_float4(ph[4], ph[5], ph[6], ph[7]);
    }
  };

  if (warp >= 8) stage_a(t_begin);
  __syncthreads();

  int par = 0;
  for (int g0 = t_begin; g0 < t_end; g0 += INV_F) {
    // (b) banded mel -> linear for magnitude and phase, then polar -> rectangular into the frame's tile
    for (;;) {
      int idx = 0;
      if (lane == 0) idx = atomicAdd(ctr, 1);
      idx = __shfl_sync(0xffffffffu, idx, 0);
      if (idx >= 32) break;
      const int d = order[idx] * 32 + lane;
      const int j0 = __ldg(pb_j0 + d);
      int cnt = __ldg(pb_cnt + d);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt = max(cnt, __shfl_xor_sync(0xffffffffu, cnt, o));
      float am[INV_F], ap[INV_F];
#pragma unroll
      for (int f = 0; f < INV_F; ++f) { am[f] = 0.0f; ap[f] = 0.0f; }
      // the band weights come from L2: four per step, the next four in flight while these are used (rows >= this
      // lane's own count hold zeros, the warp-wide maximum `cnt` only bounds the loop)
      float cn[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) cn[u] = (u < band) ? __ldg(pb_w + (size_t)u * NBINS + d) : 0.0f;
      for (int i0 = 0; i0 < cnt; i0 += 4) {
        float c4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) c4[u] = cn[u];
        if (i0 + 4 < cnt) {
#pragma unroll
          for (int u = 0; u < 4; ++u) cn[u] = (i0 + 4 + u < band) ? __ldg(pb_w + (size_t)(i0 + 4 + u) * NBINS + d) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = min(j0 + i0 + u, NBINS - 1);
          const float4* row = reinterpret_cast<const float4*>(melmp + j * MROW);
          const float4 m0 = row[0], m1 = row[1], p0 = row[2], p1 = row[3];
          const float c = c4[u];
          am[0] = fmaf(m0.x, c, am[0]); am[1] = fmaf(m0.y, c, am[1]); am[2] = fmaf(m0.z, c, am[2]); am[3] = fmaf(m0.w, c, am[3]);
          am[4] = fmaf(m1.x, c, am[4]); am[5] = fmaf(m1.y, c, am[5]); am[6] = fmaf(m1.z, c, am[6]); am[7] = fmaf(m1.w, c, am[7]);
          ap[0] = fmaf(p0.x, c, ap[0]); ap[1] = fmaf(p0.y, c, ap[1]); ap[2] = fmaf(p0.z, c, ap[2]); ap[3] = fmaf(p0.w, c, ap[3]);
          ap[4] = fmaf(p1.x, c, ap[4]); ap[5] = fmaf(p1.y, c, ap[5]); ap[6] = fmaf(p1.z, c, ap[6]); ap[7] = fmaf(p1.w, c, ap[7]);
        }
      }
#pragma unroll
      for (int f = 0; f < INV_F; ++f) {
        float sn, cs;
        sincosf(ap[f], &sn, &cs);
        tiles[f * WARP_BUF + d + 1] = am[f] * cs;
        tiles[f * WARP_BUF + XLEN + d + 1] = am[f] * sn;
      }
    }
    if (tid < INV_F) { tiles[tid * WARP_BUF] = 0.0f; tiles[tid * WARP_BUF + XLEN] = 0.0f; }
    __syncthreads();
    if (warp < INV_F) {
      // (c) one inverse real FFT per warp; the spectrum X (k = 0..1024, two planes) and the output frame share the tile
      float* tile = tiles + warp * WARP_BUF;
      const float* fxr = tile;
      const float* fxi = tile + XLEN;
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
      __syncwarp();  // every lane has read its part of X before the tile is reused for the exchange
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
      if (tid == 0) *ctr = 0;
    } else if (g0 + INV_F < t_end) {
      stage_a(g0 + INV_F);
    }
    __syncthreads();
    // (d) overlap-add: finished samples [HOP*g0, HOP*(g0+F)), ascending frame order like the reference
    {
      const float* cin = carry + par * CARRY;
      float* cout = carry + (par ^ 1) * CARRY;
      const int nf = (t_end - g0 < INV_F) ? t_end - g0 : INV_F;
      const bool head = (seg > 0) && (g0 == t_begin);   // the first 1536 samples also receive the previous segment's tail
      const int span = HOP * nf;
      // four consecutive samples per thread: they share the same (up to four) contributing frames
      const bool vec_out = ((padf & 3) == 0) && ((wave_len & 3) == 0) && ((reinterpret_cast<uintptr_t>(wave) & 15) == 0);
      for (int s = 4 * tid; s < span + CARRY; s += 4 * INV_THREADS) {
        float4 acc = (s < CARRY) ? *reinterpret_cast<const float4*>(cin + s) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int fhi = s / HOP;
        int flo = fhi - 3;
        if (flo < 0) flo = 0;
        for (int f = flo; f <= fhi && f < nf; ++f) {
          const float4 v = *reinterpret_cast<const float4*>(tiles + f * WARP_BUF + (s - HOP * f));
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        if (s < span) {
          const int ng = HOP * g0 + s - padf;
          if (head && s < CARRY) {
            if (ng >= 0 && ng < wave_len) atomicAdd(out + ng, acc.x);
            if (ng + 1 >= 0 && ng + 1 < wave_len) atomicAdd(out + ng + 1, acc.y);
            if (ng + 2 >= 0 && ng + 2 < wave_len) atomicAdd(out + ng + 2, acc.z);
            if (ng + 3 >= 0 && ng + 3 < wave_len) atomicAdd(out + ng + 3, acc.w);
          } else if (vec_out && ng >= 0 && ng + 3 < wave_len) {
            *reinterpret_cast<float4*>(out + ng) = acc;
          } else {
            if (ng >= 0 && ng < wave_len) out[ng] = acc.x;
            if (ng + 1 >= 0 && ng + 1 < wave_len) out[ng + 1] = acc.y;
            if (ng + 2 >= 0 && ng + 2 < wave_len) out[ng + 2] = acc.z;
            if (ng + 3 >= 0 && ng + 3 < wave_len) out[ng + 3] = acc.w;
          }
        } else {
          *reinterpret_cast<float4*>(cout + (s - span)) = acc;
        }
      }
      par ^= 1;
    }
    __syncthreads();
  }
  {
    const float* cin = carry + par * CARRY;
    for (int s = tid; s < CARRY; s += INV_THREADS) {
      const int ng = HOP * t_end + s - padf;
      if (ng >= 0 && ng < wave_len) {
        if (segs > 1) atomicAdd(out + ng, cin[s]);   // may overlap the next segment's head (or, for a last segment
        else out[ng] = cin[s];                       // shorter than 3 frames, the previous segment's tail)
      }
    }
  }
}

}  // namespace

extern "C" int gs_spectrogram_fwd(const float* wave, const float* hann, const int* mel_k0, const float* mel_w,
                                  float* logmel, float* inst, float* scratch, int batch, int wave_len, int time_steps,
                                  int frames_per_run, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GS_CHECK_ARG(batch >= 0 && wave_len > 0 && time_steps > 0, "spectrogram_fwd: bad shape");
  GS_CHECK_ARG(HOP * (time_steps - 1) + FRAME >= wave_len,
               "spectrogram_fwd: waveform_length %d exceeds the %d samples %d frames cover", wave_len,
               HOP * (time_steps - 1) + FRAME, time_steps);
  GS_CHECK_ARG(frames_per_run > 0, "spectrogram_fwd: frames_per_run %d must be positive", frames_per_run);
  const int runs = gs_cdiv(time_steps, frames_per_run);
  if (batch == 0) return GS_OK;
  GS_CHECK_ARG(runs == 1 || scratch != nullptr,
               "spectrogram_fwd: %d runs per clip need a scratch buffer of batch * runs * 1024 floats", runs);
  GS_NEED_CONTEXT(ctx, "spectral");
  Tables g_tables;
  int rc = ensure_tables(ctx, st, &g_tables);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    GS_CUDA(cudaFuncSetAttribute(spectrogram_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    attr = true;
  }
  const int total_runs = batch * runs;
  const int grid = total_runs < gs_num_sms() ? total_runs : gs_num_sms();      // one 16-warp CTA per SM (188 KB smem)
  spectrogram_fwd_kernel<<<grid, FWD_WARPS * 32, FWD_SMEM, st>>>(wave, wave_len, time_steps, frames_per_run, runs, hann,
                                                                g_tables.tw1024, g_tables.tw2048, mel_k0, mel_w, logmel,
                                                                inst, scratch, total_runs);
  GS_CHECK_LAUNCH("spectrogram_fwd");
  if (runs > 1) {
    const int rows = batch * (runs - 1);
    spectrogram_if_fixup_kernel<<<rows, 256, 0, st>>>(inst, scratch, time_steps, frames_per_run, runs, rows);
    GS_CHECK_LAUNCH("spectrogram_if_fixup");
  }
  return GS_OK;
}

extern "C" int gs_waveform_fwd(const float* logmel, const float* inst, const float* synth_window, const int* pb_j0,
                               const int* pb_cnt, const float* pb_w, int band, float* wave, float* scratch, int batch,
                               int wave_len, int time_steps, int frames_per_segment, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GS_CHECK_ARG(batch >= 0 && wave_len > 0 && time_steps > 0 && band > 0, "waveform_fwd: bad shape");
  GS_CHECK_ARG(HOP * (time_steps - 1) + FRAME >= wave_len,
               "waveform_fwd: waveform_length %d exceeds the %d samples %d frames cover", wave_len,
               HOP * (time_steps - 1) + FRAME, time_steps);
  GS_CHECK_ARG(frames_per_segment > 0 && (frames_per_segment >= time_steps || frames_per_segment % INV_F == 0),
               "waveform_fwd: frames_per_segment %d must be a multiple of %d (or cover all %d frames)", frames_per_segment,
               INV_F, time_steps);
  const int segs = gs_cdiv(time_steps, frames_per_segment);
  if (batch == 0) return GS_OK;
  GS_CHECK_ARG(segs == 1 || scratch != nullptr,
               "waveform_fwd: %d segments per clip need a scratch buffer of batch * segments * 1024 floats", segs);
  GS_NEED_CONTEXT(ctx, "spectral");
  Tables g_tables;
  int rc = ensure_tables(ctx, st, &g_tables);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    GS_CUDA(cudaFuncSetAttribute(waveform_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, INV_SMEM));
    attr = true;
  }
  if (segs > 1) {
    GS_CUDA(cudaMemsetAsync(wave, 0, (size_t)batch * wave_len * sizeof(float), st));
    waveform_phase_prefix_kernel<<<batch * 4, 256, 0, st>>>(inst, time_steps, frames_per_segment, segs, scratch);
    GS_CHECK_LAUNCH("waveform_phase_prefix");
  }
  waveform_fwd_kernel<<<batch * segs, INV_THREADS, INV_SMEM, st>>>(logmel, inst, time_steps, wave_len, synth_window,
                                                                  g_tables.tw1024, g_tables.tw2048, pb_j0, pb_cnt, pb_w,
                                                                  band, wave, frames_per_segment, segs, scratch);
  GS_CHECK_LAUNCH("waveform_fwd");
  return GS_OK;
}
