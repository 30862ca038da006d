// Generic spectral front-end: any number of frequency bins and any overlap (reference spectral_ops.py:45-149 is
// generic in spectrogram_shape / overlap; BASELINE config 1 uses a [16, 16] spectrogram: frame 32, hop 8).
// spectral.cu holds the kernels specialised to the reference's production configuration (1024 bins, 75 % overlap);
// everything else runs here: direct DFTs and dense mel / pseudo-inverse products, one CTA per frame.  Small
// configurations are launch-bound, large ones O(bins^2) per frame -- a correct general path, not a tuned one.
#include "common.cuh"
#include "gansynth_b200.h"

namespace {

constexpr float PI_F = 3.14159265358979323846f;

// ---- forward, pass 1: frame -> |DFT|, arg(DFT) (bins 1..bins) -> mel products -> log-mel (scaled) and mel phase ------
// grid (T, batch); dynamic shared memory: frame [2 bins] | mag [bins] | phase [bins]
__global__ void spectrogram_generic_kernel(const float* __restrict__ wave, const float* __restrict__ hann,
                                           const float* __restrict__ mel /*[bins][bins]*/, float* __restrict__ logmel,
                                           float* __restrict__ melphase, int wave_len, int time_steps, int bins, int step,
                                           int pad) {
  extern __shared__ float sm[];
  const int n_fft = 2 * bins;
  float* frame = sm;
  float* mag = sm + n_fft;
  float* phase = mag + bins;
  const int t = blockIdx.x, b = blockIdx.y;
  for (int n = threadIdx.x; n < n_fft; n += blockDim.x) {
    const int src = t * step + n - pad;                 // the clip is left-padded with `pad` zeros (spectral_ops.py:55-57)
    frame[n] = (src >= 0 && src < wave_len) ? wave[(size_t)b * wave_len + src] * hann[n] : 0.0f;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < bins; k += blockDim.x) {
    const int kk = k + 1;                               // the DC bin is dropped (spectral_ops.py:66)
    float re = 0.0f, im = 0.0f;
    for (int n = 0; n < n_fft; ++n) {
      float s, c;
      sincospif(2.0f * (float)((kk * n) % n_fft) / (float)n_fft, &s, &c);
      re = fmaf(frame[n], c, re);
      im = fmaf(-frame[n], s, im);
    }
    mag[k] = sqrtf(re * re + im * im);
    phase[k] = atan2f(im + 0.0f, re + 0.0f);            // + 0: silence has phase 0 whatever the signed zeros are
  }
  __syncthreads();
  for (int j = threadIdx.x; j < bins; j += blockDim.x) {
    float mm = 0.0f, mp = 0.0f;
    for (int k = 0; k < bins; ++k) {
      const float w = mel[(size_t)k * bins + j];
      mm = fmaf(mag[k], w, mm);
      mp = fmaf(phase[k], w, mp);
    }
    const size_t o = ((size_t)b * time_steps + t) * bins + j;
    logmel[o] = (logf(mm + 1.0e-6f) - (-3.76f)) / 10.05f;
    melphase[o] = mp;
  }
}

// floor-mod phase wrap of spectral_ops.py:20-31 applied to one difference
__device__ __forceinline__ float wrap_diff(float d) {
  const float two_pi = 2.0f * PI_F;
  float m = d + PI_F;
  m = m - two_pi * floorf(m / two_pi) - PI_F;
  if (m == -PI_F && d > 0.0f) m = PI_F;
  return m;
}

// ---- forward, pass 2: instantaneous frequency = unwrap -> diff -> / pi, i.e. wrap(phase[t] - phase[t-1]) / pi ---------
__global__ void if_generic_kernel(const float* __restrict__ melphase, float* __restrict__ inst, int time_steps, int bins,
                                  size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)((i / bins) % time_steps);
    inst[i] = (t == 0 ? melphase[i] : wrap_diff(melphase[i] - melphase[i - bins])) / PI_F;
  }
}

// ---- inverse, pass 1: mel phase = cumulative sum over time of pi * IF (spectral_ops.py:112-113) ---------------------
__global__ void phase_cumsum_generic_kernel(const float* __restrict__ inst, float* __restrict__ melphase, int batch,
                                            int time_steps, int bins) {
  const size_t total = (size_t)batch * bins;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / bins, j = i % bins;
    float acc = 0.0f;
    for (int t = 0; t < time_steps; ++t) {
      const size_t o = (b * time_steps + t) * bins + j;
      acc += inst[o] * PI_F;
      melphase[o] = acc;
    }
  }
}

// ---- inverse, pass 2: mel -> linear (pseudo-inverse), polar -> rectangular, inverse real DFT, synthesis window -------
// grid (T, batch); shared: melmag [bins] | melphase [bins] | re [bins] | im [bins]; frames [batch, T, 2 bins]
__global__ void waveform_frames_generic_kernel(const float* __restrict__ logmel, const float* __restrict__ melphase,
                                               const float* __restrict__ pinv /*[bins mel][bins linear]*/,
                                               const float* __restrict__ synth, float* __restrict__ frames, int time_steps,
                                               int bins) {
  extern __shared__ float sm[];
  const int n_fft = 2 * bins;
  float* mm = sm;
  float* mp = mm + bins;
  float* re = mp + bins;
  float* im = re + bins;
  const int t = blockIdx.x, b = blockIdx.y;
  const size_t row = ((size_t)b * time_steps + t) * bins;
  for (int j = threadIdx.x; j < bins; j += blockDim.x) {
    mm[j] = expf(logmel[row + j] * 10.05f + (-3.76f));
    mp[j] = melphase[row + j];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < bins; k += blockDim.x) {
    float a = 0.0f, p = 0.0f;
    for (int j = 0; j < bins; ++j) {
      const float w = pinv[(size_t)j * bins + k];
      a = fmaf(mm[j], w, a);
      p = fmaf(mp[j], w, p);
    }
    float s, c;
    sincosf(p, &s, &c);
    re[k] = a * c;                                      // rfft bin k + 1 (bin 0 is the zero the reference pads back in)
    im[k] = a * s;
  }
  __syncthreads();
  const float inv_n = 1.0f / (float)n_fft;
  for (int n = threadIdx.x; n < n_fft; n += blockDim.x) {
    // irfft: x[n] = (1/N) (X0 + 2 sum_{k=1}^{N/2-1} Re(X_k e^{2 pi i k n / N}) + Re(X_{N/2}) (-1)^n), X0 = 0
    float acc = 0.0f;
    for (int k = 1; k < bins; ++k) {
      float s, c;
      sincospif(2.0f * (float)((k * n) % n_fft) / (float)n_fft, &s, &c);
      acc += re[k - 1] * c - im[k - 1] * s;
    }
    acc = 2.0f * acc + ((n & 1) ? -re[bins - 1] : re[bins - 1]);
    frames[(((size_t)b * time_steps + t) * n_fft) + n] = acc * inv_n * synth[n];
  }
}

// ---- inverse, pass 3: overlap-add (gather form) and the crop of the left padding ------------------------------------
__global__ void overlap_add_generic_kernel(const float* __restrict__ frames, float* __restrict__ wave, int batch, int wave_len,
                                           int time_steps, int bins, int step, int pad) {
  const int n_fft = 2 * bins;
  const size_t total = (size_t)batch * wave_len;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / wave_len;
    const int pos = (int)(i % wave_len) + pad;          // position in the padded signal
    int t_hi = pos / step;
    if (t_hi > time_steps - 1) t_hi = time_steps - 1;
    float acc = 0.0f;
    for (int t = t_hi; t >= 0 && pos - t * step < n_fft; --t)
      acc += frames[((b * time_steps + t) * n_fft) + (pos - t * step)];
    wave[i] = acc;
  }
}

}  // namespace

#define ST ((cudaStream_t)stream)

static int generic_check(const char* who, int batch, int wave_len, int time_steps, int bins, int step) {
  GS_CHECK_ARG(batch >= 0 && wave_len > 0 && time_steps > 0 && bins > 0 && step > 0 && step <= 2 * bins,
               "%s: bad shape (bins %d, step %d)", who, bins, step);
  GS_CHECK_ARG(step * (time_steps - 1) + 2 * bins >= wave_len, "%s: waveform_length %d exceeds the %d samples %d frames cover",
               who, wave_len, step * (time_steps - 1) + 2 * bins, time_steps);
  GS_CHECK_ARG((size_t)6 * bins * sizeof(float) <= 200 * 1024, "%s: %d bins exceed the shared-memory frame buffers", who, bins);
  return GS_OK;
}

// spectral_ops.py:45-94 for any (bins, frame_step): hann [2 bins], mel dense [bins][bins] (linear x mel, DC row dropped),
// scratch = caller-owned [batch, T, bins] floats (mel phases)
extern "C" int gs_spectrogram_generic(const float* wave, const float* hann, const float* mel, float* logmel, float* inst,
                                      float* scratch, int batch, int wave_len, int time_steps, int bins, int frame_step,
                                      void* stream) {
  int rc = generic_check("spectrogram_generic", batch, wave_len, time_steps, bins, frame_step);
  if (rc) return rc;
  if (batch == 0) return GS_OK;
  GS_CHECK_ARG(scratch != nullptr, "spectrogram_generic: scratch of batch * T * bins floats is required");
  const int pad = frame_step * (time_steps - 1) + 2 * bins - wave_len;
  const size_t smem = (size_t)4 * bins * sizeof(float);
  if (smem > 48 * 1024) GS_CUDA(cudaFuncSetAttribute(spectrogram_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = bins >= 256 ? 256 : bins >= 64 ? 128 : 64;
  spectrogram_generic_kernel<<<dim3((unsigned)time_steps, (unsigned)batch), threads, smem, ST>>>(wave, hann, mel, logmel, scratch, wave_len,
                                                                                                 time_steps, bins, frame_step, pad);
  GS_CHECK_LAUNCH("spectrogram_generic");
  const size_t total = (size_t)batch * time_steps * bins;
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)gs_num_sms() * 16;
  if_generic_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ST>>>(scratch, inst, time_steps, bins, total);
  GS_CHECK_LAUNCH("if_generic");
  return GS_OK;
}

// spectral_ops.py:97-149 for any (bins, frame_step): synth_window [2 bins], pinv dense [bins mel][bins linear];
// scratch = caller-owned [batch, T, 3 bins] floats (mel phases, then the windowed frames)
extern "C" int gs_waveform_generic(const float* logmel, const float* inst, const float* synth_window, const float* pinv,
                                   float* wave, float* scratch, int batch, int wave_len, int time_steps, int bins,
                                   int frame_step, void* stream) {
  int rc = generic_check("waveform_generic", batch, wave_len, time_steps, bins, frame_step);
  if (rc) return rc;
  if (batch == 0) return GS_OK;
  GS_CHECK_ARG(scratch != nullptr, "waveform_generic: scratch of batch * T * 3 * bins floats is required");
  float* melphase = scratch;
  float* frames = scratch + (size_t)batch * time_steps * bins;
  const int pad = frame_step * (time_steps - 1) + 2 * bins - wave_len;
  const size_t cap = (size_t)gs_num_sms() * 16;
  size_t blocks = ((size_t)batch * bins + 127) / 128;
  phase_cumsum_generic_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 128, 0, ST>>>(inst, melphase, batch, time_steps, bins);
  GS_CHECK_LAUNCH("phase_cumsum_generic");
  const size_t smem = (size_t)4 * bins * sizeof(float);
  if (smem > 48 * 1024) GS_CUDA(cudaFuncSetAttribute(waveform_frames_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = bins >= 128 ? 256 : bins >= 32 ? 128 : 64;
  waveform_frames_generic_kernel<<<dim3((unsigned)time_steps, (unsigned)batch), threads, smem, ST>>>(logmel, melphase, pinv, synth_window,
                                                                                                     frames, time_steps, bins);
  GS_CHECK_LAUNCH("waveform_frames_generic");
  blocks = ((size_t)batch * wave_len + 255) / 256;
  overlap_add_generic_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ST>>>(frames, wave, batch, wave_len, time_steps, bins,
                                                                                      frame_step, pad);
  GS_CHECK_LAUNCH("overlap_add_generic");
  return GS_OK;
}
