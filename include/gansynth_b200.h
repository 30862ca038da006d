/* gansynth_b200 -- C ABI of the B200-native GANSynth hot path (libgansynth_b200.so, sm_100a).
 *
 * The reference (skmhrk1209/GANSynth @ d135d40) has no FFI: its hot path is TensorFlow-1.13 library
 * calls made from ops.py / spectral_ops.py / models.py.  Each entry point below replaces one of those
 * call sites (file:line given per function) and is what a host binding (ctypes here, see
 * gansynth_b200/_lib.py and INTEGRATION.md) binds.
 *
 * Conventions
 *  - every pointer is a caller-owned DEVICE pointer to contiguous fp32 (int32 / int64 where stated);
 *    the library never frees or retains them past the call;
 *  - activations are NHWC ([n, h, w, c], c fastest); conv weights keep the TF variable layout
 *    [kh, kw, cin, cout] (`wswap`=1: the tensor in memory is [kh, kw, cout, cin]);
 *  - `alpha` is the run-time equalised-learning-rate constant of get_weight (ops.py:154-160), applied
 *    to the contraction result (before bias);
 *  - `act`: 0 none, 1 leaky-relu(0.2) applied after the bias;
 *  - `impl`: 0 auto (tensor-core kernel where the shape allows, else fp32), 1 naive anchor kernel,
 *    2 tiled fp32 FFMA kernel, 3 tcgen05 tensor-core kernel (2/3: error if the shape is unsupported),
 *    4 fp32 auto (tiled else naive; never the tensor-core kernel); OR-ed with GS_IMPL_PARAM_WEIGHT when `w` is a
 *    network parameter (a pointer that stays valid and unchanged until gs_conv_weight_cache_reset): the
 *    tensor-core kernels may then reuse its pre-split bf16 copy across calls;
 *  - `stream` is a cudaStream_t; every call is asynchronous on it: no entry point synchronises, allocates or frees
 *    device memory, or copies from the host;
 *  - the library keeps NO global mutable state.  What survives a call -- the bf16-split copies of parameter weights the
 *    tensor-core convolutions reuse within a sub-step, and the spectral twiddle tables -- lives in a caller-owned
 *    device workspace behind an opaque gs_context (below), bound per host thread;
 *  - return 0 on success, negative on error; gs_last_error() gives the message (thread-local).
 */
#ifndef GANSYNTH_B200_H_
#define GANSYNTH_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* gs_last_error(void);
int gs_version(void);

/* ---- context: caller-owned workspace ------------------------------------------------------------------------------
 * gs_workspace_bytes(): the recommended size of the device workspace (twiddle tables + an 8 MB slot for split weights that
 * are used once + a 256 MB cache for split parameter weights); gs_workspace_min_bytes(): the smallest accepted (no
 * cache: every parameter is re-split at each use).  gs_context_create() only records the pointer (256-byte aligned;
 * it stays the caller's: the library never frees it) -- no device work.  gs_context_bind() makes `ctx` the context of the
 * CALLING HOST THREAD (NULL unbinds); the convolution, spectral and cache-reset entry points use the bound context and
 * fail with GS_ERR_ARG when there is none.  A context serves one stream at a time: the caller orders hand-overs
 * between streams as for any other buffer (the twiddle tables are filled by a kernel enqueued on the stream of the
 * context's first spectral call).  A full cache degrades to re-splitting, with one warning on stderr. */
typedef struct gs_context gs_context;
size_t gs_workspace_bytes(void);
size_t gs_workspace_min_bytes(void);
int gs_context_create(void* workspace, size_t bytes, gs_context** out);
int gs_context_destroy(gs_context* ctx);
int gs_context_bind(gs_context* ctx);

#define GS_IMPL_PARAM_WEIGHT 0x100
/* Invalidates the bound context's cache of pre-split parameter weights.  Call after every optimiser update and at the start of
   every sub-step (TF evaluates get_weight's scaling at run time, ops.py:154-160: nothing may survive an update). */
int gs_conv_weight_cache_reset(void);
/* Re-splits in place every cached parameter weight inside [lo, hi) (NULL, NULL: all) after the caller changed those
   parameters: the cache slots keep their addresses (kernels recorded in CUDA graphs stay valid), one launch per 48
   entries.  With a refresh after every change the cache never holds a stale copy and a sub-step contains no split kernels. */
int gs_conv_weight_cache_refresh(const float* lo, const float* hi, void* stream);

/* ---- convolution family: tf.nn.conv2d ops.py:237-243 (+ bias_add :245-246) and its gradients ------
 * SAME padding as TF computes it for even sizes: 3x3 stride 1 pads (1,1); stride 2 pads (0,1).
 * fwd  : y[n,h/s,w/s,co]  = act(alpha * sum x[n, oh*s+kh-pb, ow*s+kw-pb, ci] * W(kh,kw,ci,co) + bias[co])
 * dgrad: dx[n,h,w,ci]     = act(alpha * sum dy[n,oh,ow,co] * W(kh,kw,ci,co) + bias[ci])   (oh*s+kh-pb = ih)
 * wgrad: dw(kh,kw,ci,co)  = alpha * sum x[n, oh*s+kh-pb, ow*s+kw-pb, ci] * dy[n,oh,ow,co]
 * h, w are always the spatial size of the LARGE side (x of fwd / dx of dgrad). bias may be NULL. */
int gs_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, int n, int h, int wd, int ci, int co,
                  int ksize, int stride, int wswap, float alpha, int act, int impl, void* stream);
int gs_conv2d_dgrad(const float* dy, const float* w, const float* bias, float* dx, int n, int h, int wd, int ci,
                    int co, int ksize, int stride, int wswap, float alpha, int act, int impl, void* stream);
int gs_conv2d_wgrad(const float* x, const float* dy, float* dw, int n, int h, int wd, int ci, int co, int ksize,
                    int stride, int wswap, float alpha, int impl, void* stream);
/* gs_conv2d_wgrad plus, from the same pass over the operands, the bias gradient (BiasAddGrad of ops.py:244 / 277):
 * dbias[c] = sum over pixels of dy (bias_of_x = 0, conv2d layers) or of x (bias_of_x = 1: conv2d_transpose layers,
 * whose pre-activation gradient is the high-resolution operand `x` of this call); dbias ([co] / [ci]) may be NULL.
 * accumulate = 0 overwrites dw / dbias; 1 ADDS to them: the uses of one parameter (D(real), D(fake), the penalty pass)
 * then sum straight into the caller's gradient buffer -- tf.gradients' AddN (models.py:81-89) without extra passes. */
int gs_conv2d_wgrad_ex(const float* x, const float* dy, float* dw, float* dbias, int bias_of_x, int accumulate, int n,
                       int h, int wd, int ci, int co, int ksize, int stride, int wswap, float alpha, int impl,
                       void* stream);

/* The same two convolutions with a FUSED EPILOGUE, `epi`:
 *   GS_EPI_NONE  (0): as above.
 *   GS_EPI_MASK  (1): out = lrelu'(aux) * alpha * conv(...).  `aux` is shaped like the output; only its sign is used.  This
 *                     is tf.nn.leaky_relu's gradient (LeakyReluGrad of networks.py:184,194,216,227: the mask of the layer
 *                     whose output the convolution's result is a gradient of) applied while the tile is still on chip.
 *                     bias must be NULL, act 0.
 *   GS_EPI_PIXEL_NORM (2): out = pixel_normalization(leaky_relu(alpha * conv + bias)) over the output channels (ops.py:330-333
 *                     after networks.py:57-68, 82-93), rvec[n, oh, ow] = 1 / sqrt(mean_c(a^2) + eps).  act must be 1.
 * The result is always the fused one; shapes the tensor-core epilogue does not cover run the plain convolution followed by
 * the elementwise kernel, in place. */
#define GS_EPI_NONE 0
#define GS_EPI_MASK 1
#define GS_EPI_PIXEL_NORM 2
int gs_conv2d_fwd_ex(const float* x, const float* w, const float* bias, float* y, int n, int h, int wd, int ci, int co,
                     int ksize, int stride, int wswap, float alpha, int act, int epi, const float* aux, float* rvec,
                     float eps, int impl, void* stream);
int gs_conv2d_dgrad_ex(const float* dy, const float* w, const float* bias, float* dx, int n, int h, int wd, int ci,
                       int co, int ksize, int stride, int wswap, float alpha, int act, int epi, const float* aux,
                       float* rvec, float eps, int impl, void* stream);

/* ---- tf.nn.conv2d_transpose ops.py:266-276: x [n,h,w,cin], var [k,k,cin,filters] -> y [n,h*s,w*s,filters] */
int gs_conv2d_transpose_fwd(const float* x, const float* var, const float* bias, float* y, int n, int h, int wd,
                            int cin, int filters, int ksize, int stride, float alpha, int act, int impl, void* stream);
int gs_conv2d_transpose_dgrad(const float* dy, const float* var, float* dx, int n, int h, int wd, int cin, int filters,
                              int ksize, int stride, float alpha, int impl, void* stream);
int gs_conv2d_transpose_wgrad(const float* x, const float* dy, float* dvar, int n, int h, int wd, int cin, int filters,
                              int ksize, int stride, float alpha, int impl, void* stream);

/* ---- dense: tf.matmul ops.py:197 and its gradients; w is [k, n] ---------------------------------- */
int gs_dense_fwd(const float* x, const float* w, float* y, int m, int k, int n, float alpha, void* stream);
int gs_dense_dgrad(const float* dy, const float* w, float* dx, int m, int k, int n, float alpha, void* stream);
int gs_dense_wgrad(const float* x, const float* dy, float* dw, int m, int k, int n, float alpha, void* stream);

/* ---- embedding: tf.nn.embedding_lookup ops.py:217; idx is int64 [b] (argmax of the one-hot labels) - */
int gs_embedding_fwd(const float* table, const long long* idx, float* out, int b, int units, float alpha, void* stream);
int gs_embedding_bwd(const float* dy, const long long* idx, float* dtable, int b, int rows, int units, float alpha,
                     void* stream);

/* ---- activations (networks.py: tf.nn.leaky_relu, tf.nn.tanh) and their derivative forms ---------- */
int gs_lrelu(const float* x, float* out, long long n, void* stream);
int gs_lrelu_mask_mul(const float* v, const float* y, float* out, long long n, void* stream); /* v * (y>0 ? 1 : 0.2) */
int gs_tanh_fwd(const float* x, float* out, long long n, void* stream);
int gs_tanh_bwd(const float* y, const float* dy, float* out, long long n, void* stream);        /* dy (1 - y^2) */
int gs_tanh_bwd2(const float* y, const float* dy, const float* u, float* out, long long n, void* stream); /* -2 y dy u */
int gs_bias_act(const float* x, const float* bias, float* out, long long rows, int c, int act, void* stream);
int gs_row_broadcast(const float* s, float* out, long long rows, int c, void* stream);
int gs_col_sum(const float* v, float* out, long long rows, int c, void* stream);                /* bias gradient */
/* mask-multiply and the bias gradient of the same layer in one pass (the un-fused pair above is the TF graph's
   LeakyReluGrad + BiasAddGrad of ops.py:237-247 / networks.py); c % 4 == 0, c <= 256, c/4 divides 256 */
int gs_lrelu_mask_mul_colsum(const float* v, const float* y, float* out, float* colsum, long long rows, int c, void* stream);

/* ---- lerp networks.py:10-11 and generic linear combinations ------------------------------------- */
int gs_axpby(const float* a, const float* b, float* out, float alpha, float beta, long long n, void* stream);
/* lerp (networks.py:10-11) with the blend weights in DEVICE memory: out = coef[ia] * a + coef[ib] * b (b may be NULL).
   Lets the progressive-growing sub-steps be replayed as CUDA graphs while the weight follows global_step. */
int gs_axpby_dev(const float* a, const float* b, float* out, const float* coef, int ia, int ib, long long n,
                 void* stream);
int gs_mul(const float* a, const float* b, float* out, float alpha, long long n, void* stream);

/* ---- pixel_normalization ops.py:330-333 over the channel axis of [rows, c] ----------------------- */
int gs_pixel_norm_fwd(const float* a, float* y, float* r, long long rows, int c, float eps, void* stream);
int gs_pixel_norm_bwd(const float* a, const float* r, const float* dy, float* da, long long rows, int c, void* stream);
/* pixel-norm backward fused with the leaky-relu mask of the layer that produced a (networks.py:57-68,82-93: conv ->
   leaky_relu -> pixel_normalization); colsum (may be null) receives the bias gradient; c % 4 == 0, c <= 256 */
int gs_pixel_norm_bwd_mask(const float* a, const float* r, const float* dy, float* dz, float* colsum, long long rows, int c,
                           void* stream);
/* second-order forms of gs_pixel_norm_bwd_mask (M = lrelu'(a)): J(a)(M u) and M * bwd2(a, r, dy, M u) */
int gs_pixel_norm_bwd_premask(const float* a, const float* r, const float* u, float* out, long long rows, int c, void* stream);
int gs_pixel_norm_bwd2_masked(const float* a, const float* r, const float* dy, const float* u, float* ga, long long rows, int c,
                              void* stream);
int gs_pixel_norm_bwd2(const float* a, const float* r, const float* dy, const float* u, float* ga, long long rows,
                       int c, void* stream);
/* "y form" of the fused gradients above for layers run with GS_EPI_PIXEL_NORM, which keep only the normalised output
   y = a * r and r (a = y / r is rebuilt in registers): networks.py:57-68, 82-93 differentiated */
int gs_pixel_norm_bwd_mask_y(const float* y, const float* r, const float* dy, float* dz, float* colsum, long long rows, int c,
                             void* stream);
int gs_pixel_norm_bwd_premask_y(const float* y, const float* r, const float* u, float* out, long long rows, int c, void* stream);
int gs_pixel_norm_bwd2_masked_y(const float* y, const float* r, const float* dy, const float* u, float* ga, long long rows, int c,
                                void* stream);
/* both second-order pieces for one incoming u in a single pass: ga as gs_pixel_norm_bwd2_masked_y, gdy as
   gs_pixel_norm_bwd_premask_y(y, r, u) */
int gs_pixel_norm_bwd2_pair_y(const float* y, const float* r, const float* dy, const float* u, float* ga, float* gdy,
                              long long rows, int c, void* stream);

/* ---- batch_stddev ops.py:336-348 on [b, e]; stat is [b/groups] ---------------------------------- */
int gs_batch_stddev_fwd(const float* x, float* stat, int b, long long e, int groups, float eps, void* stream);
int gs_batch_stddev_bwd(const float* x, const float* df, float* dx, int b, long long e, int groups, float eps,
                        void* stream);
int gs_batch_stddev_bwd2(const float* x, const float* df, const float* u, float* gx, float* q, int b, long long e,
                         int groups, float eps, void* stream);

/* ---- upscale2d ops.py:283-291 / downscale2d ops.py:294-305 (NHWC), layout permutation ------------ */
int gs_upscale2d(const float* in, float* out, int n, int h, int w, int c, int fh, int fw, float scale, void* stream);
int gs_pool2d(const float* in, float* out, int n, int h, int w, int c, int fh, int fw, float scale, void* stream);
int gs_transpose_inner(const float* in, float* out, int n, int a, int b, void* stream); /* [n,a,b] -> [n,b,a] */

/* ---- per-sample reductions of the penalties models.py:48,61 on [rows, e] ------------------------- */
int gs_row_dot(const float* a, const float* b, float* out, int rows, long long e, void* stream);
int gs_row_scale(const float* a, const float* s, float* out, int rows, long long e, float alpha, void* stream);

/* ---- tf.train.AdamOptimizer models.py:67-76 (epsilon outside the bias correction), t = 1-based step */
int gs_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                 float eps, long long t, float grad_scale, void* stream);

/* ---- data-parallel update: gradient all-reduce FUSED with Adam over NVLink / NVSwitch (models.py:81-89 on W ranks) ------
 * One kernel per network instead of ncclAllReduce + gs_adam_step.  The flat gradient and parameter buffers live in
 * SYMMETRIC memory (the same allocation mapped on every rank; the host obtains the mappings, e.g. through
 * torch.distributed._symmetric_memory).  Rank r owns the slice gs_adam_slice(n, r, W): it reads the sum of all ranks'
 * gradients of that slice -- `grad_multicast` != NULL: multimem.ld_reduce on the NVSwitch multicast address (the switch
 * adds); else plain loads through `grad_peers[W]` (device array of the peers' pointers) -- applies TF-Adam to the slice
 * (m, v are therefore SHARDED: only the owner's slice is valid) and writes the new parameters into every rank's buffer
 * (multimem.st on `param_multicast`, else stores through `param_peers[W]`).  `p_local`: this rank's parameter buffer.
 * The CALLER brackets the launch with cross-rank barriers: every rank's gradients complete before, every rank's
 * parameters visible after.  grad_scale = 1 / W for mean-reduced losses. */
int gs_adam_slice(long long n, int rank, int world, long long* lo, long long* hi);
int gs_adam_step_allreduce(const float* p_local, float* m, float* v, const float* grad_multicast, float* param_multicast,
                           const float* const* grad_peers, float* const* param_peers, long long n, int rank, int world,
                           float lr, float beta1, float beta2, float eps, long long t, float grad_scale, void* stream);

/* ---- spectral front-end, reference configuration frame 2048 / hop 512 / 1024 bins ---------------
 * gs_spectrogram_fwd: spectral_ops.py:45-94.  wave [batch, wave_len] -> logmel, inst [batch, T, 1024].
 *   hann [2048]; mel_k0 int32 [1024] and mel_w [6][1024]: column-sparse linear->mel matrix (first
 *   non-zero row and up to 6 weights per mel bin).  One CTA walks a run of frames_per_run consecutive
 *   frames of a clip; scratch: caller-owned [batch, ceil(T / frames_per_run), 1024] floats (run-boundary
 *   phases; may be NULL when frames_per_run >= T).
 * gs_waveform_fwd: spectral_ops.py:97-149.  logmel, inst -> wave [batch, wave_len].
 *   synth_window [2048] (hann / overlap-added hann^2); pb_j0/pb_cnt int32 [1024], pb_w [band][1024]:
 *   banded pseudo-inverse (first mel row, row count and zero-padded weights per linear bin).
 *   frames_per_segment >= T: one CTA per clip.  Smaller (a multiple of 8; for batches that do not fill the
 *   GPU): one CTA per segment, scratch = caller-owned [batch, ceil(T / frames_per_segment), 1024] floats
 *   (phase prefix per segment). */
int gs_spectrogram_fwd(const float* wave, const float* hann, const int* mel_k0, const float* mel_w, float* logmel,
                       float* inst, float* scratch, int batch, int wave_len, int time_steps, int frames_per_run,
                       void* stream);
int gs_waveform_fwd(const float* logmel, const float* inst, const float* synth_window, const int* pb_j0,
                    const int* pb_cnt, const float* pb_w, int band, float* wave, float* scratch, int batch,
                    int wave_len, int time_steps, int frames_per_segment, void* stream);

/* ---- input pipeline natives, reference dataset.py:12-91 (tf.data.TFRecordDataset, tf.read_file,
 * audio_ops.decode_wav).  gs_crc32c / gs_wav_decode_pcm16 / gs_wav_read_batch are HOST functions on HOST
 * pointers; gs_pcm16_to_float is the device half of decode_wav (int16 -> float32 / 32768).
 * gs_crc32c: CRC-32C (Castagnoli) of n bytes -- the TFRecord framing checksum (before TF's mask rotation).
 * gs_wav_decode_pcm16: RIFF/WAVE 16-bit PCM bytes -> channel 0, cropped / zero-padded at the end to
 *   desired_samples int16 (dataset.py:32-36: desired_channels=1, desired_samples=64000).
 * gs_wav_read_batch: reads and decodes n files on `threads` host threads into dst [n, desired_samples]
 *   (caller-owned, normally pinned); status [n] (may be NULL) receives 0 or a negative code per file. */
int gs_crc32c(const void* data, long long n, unsigned int* out);
int gs_wav_decode_pcm16(const void* file_bytes, long long n, short* dst, int desired_samples, int* sample_rate,
                        int* samples_in_file);
int gs_wav_read_batch(const char* const* paths, int n, short* dst, int desired_samples, int threads, int* status);
int gs_pcm16_to_float(const short* src, float* dst, long long n, void* stream);

/* ---- spectral front-end, ANY configuration (spectral_ops.py:50-53 is generic in spectrogram_shape / overlap; BASELINE
 * config 1 uses a [16, 16] spectrogram: frame 32, hop 8): direct DFTs and dense mel / pseudo-inverse products, one CTA per
 * frame.  hann, synth_window [2 bins]; mel [bins][bins] (linear x mel, DC row dropped), pinv [bins mel][bins linear];
 * frame_step = int((1 - overlap) * 2 bins).  scratch is caller-owned: batch * T * bins floats (forward), batch * T * 3 bins
 * floats (inverse). */
int gs_spectrogram_generic(const float* wave, const float* hann, const float* mel, float* logmel, float* inst, float* scratch,
                           int batch, int wave_len, int time_steps, int bins, int frame_step, void* stream);
int gs_waveform_generic(const float* logmel, const float* inst, const float* synth_window, const float* pinv, float* wave,
                        float* scratch, int batch, int wave_len, int time_steps, int bins, int frame_step, void* stream);

/* ---- ResNet pitch classifier (networks.py:293-413; features for evaluate, models.py:196-230; trained by models.py:253-410)
 * group_normalization ops.py:118-146 on NHWC [n, hw, c]: per (sample, group) mean / biased variance over (hw, c/groups),
 * y = (x - mean) / sqrt(var + eps) * gamma[c] + beta[c], optionally followed by tf.nn.relu (networks.py:318-322);
 * `stats` is caller scratch of n * groups * 2 floats.  max_pooling2d ops.py:308-316 (TF SAME).  spatial_mean:
 * tf.reduce_mean over the image axes (networks.py:399) -> [n, c].  The convolutions use the family above (7x7 / 5x5
 * kernels and 1x1 stride 2 run on the generic kernels). */
int gs_group_norm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* stats, int n, long long hw,
                      int c, int groups, float eps, int relu, void* stream);
int gs_max_pool2d(const float* x, float* y, int n, int h, int w, int c, int ksize, int stride, void* stream);
int gs_spatial_mean(const float* x, float* y, int n, long long hw, int c, void* stream);
/* Gradients for TRAINING the classifier (models.py:253-304).  gs_group_norm_bwd: x, y (forward output: the relu mask),
 * dy, stats (as written by gs_group_norm_fwd) -> dx, dgamma [c], dbeta [c]; red = caller scratch of n * groups * 2 floats.
 * gs_max_pool2d_bwd: a window's gradient goes to its first maximum in row-major order, as tf's MaxPoolGrad (it matters on
 * exact ties: constant regions).  gs_momentum_step: tf.train.MomentumOptimizer
 * (models.py:283-295) on flat buffers with the L2 weight decay of models.py:267-271 folded in as wd[i] * p[i] (wd NULL or a
 * per-element coefficient, 0 on the normalisation variables):  accum = momentum * accum + g;
 * p -= lr * (nesterov ? g + momentum * accum : accum). */
int gs_group_norm_bwd(const float* x, const float* y, const float* dy, const float* stats, const float* gamma, float* dx,
                      float* dgamma, float* dbeta, float* red, int n, long long hw, int c, int groups, float eps, int relu,
                      void* stream);
int gs_max_pool2d_bwd(const float* x, const float* y, const float* dy, float* dx, int n, int h, int w, int c, int ksize,
                      int stride, void* stream);
int gs_spatial_mean_bwd(const float* dy, float* dx, int n, long long hw, int c, void* stream);
int gs_momentum_step(float* p, const float* g, float* accum, const float* wd, long long n, float lr, float momentum,
                     int nesterov, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GANSYNTH_B200_H_ */
