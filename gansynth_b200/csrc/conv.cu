// C-ABI entry points for the NHWC convolution family (see include/gansynth_b200.h).
// Replaces the tf.nn.conv2d / tf.nn.conv2d_transpose call sites of the reference
// (ops.py:237, ops.py:269) and their TF-generated gradients (models.py:47,60,81-89).
#include <stdlib.h>
#include <algorithm>

#include "conv1x1.cuh"
#include "conv_tc.cuh"
#include "conv_tck.cuh"
#include "conv_tcw.cuh"
#include "conv_tiled.cuh"
#include "tma_host.h"
#include "gansynth_b200.h"

// elementwise fallbacks of the fused epilogues (elementwise.cu)
extern "C" int gs_lrelu_mask_mul(const float* v, const float* y, float* out, long long n, void* stream);
extern "C" int gs_pixel_norm_fwd(const float* a, float* y, float* r, long long rows, int c, float eps, void* stream);

namespace {

// Fused epilogue request of the *_ex entry points (TcEpi of conv_tc.cuh)
struct Epi {
  int mode = 0;
  const float* aux = nullptr;     // MASK: mask source, shaped like the output
  float* rvec = nullptr;          // PNF: per-pixel 1 / sqrt(mean_c(a^2) + eps)
  float eps = 0.0f;
};

int make_geom(ConvGeom& g, int n, int h, int w, int ci, int co, int ksize, int stride, int wswap, float alpha,
              int act) {
  GS_CHECK_ARG(n > 0 && h > 0 && w > 0 && ci > 0 && co > 0, "conv: non-positive dimension");
  GS_CHECK_ARG(ksize == 1 || ksize == 3 || ksize == 5 || ksize == 7, "conv: ksize must be 1, 3, 5 or 7 (got %d)", ksize);
  GS_CHECK_ARG(stride == 1 || stride == 2, "conv: stride must be 1 or 2 (got %d)", stride);
  GS_CHECK_ARG(h % stride == 0 && w % stride == 0, "conv: spatial size %dx%d not divisible by stride %d", h, w, stride);
  GS_CHECK_ARG(act == 0 || act == 1, "conv: act must be 0 (none) or 1 (leaky relu)");
  g.n = n; g.h = h; g.w = w; g.ci = ci; g.co = co;
  g.oh = h / stride; g.ow = w / stride;
  g.ksize = ksize; g.stride = stride;
  // TF SAME on sizes divisible by the stride: total padding max(ksize - stride, 0), the smaller half in front
  // (3x3: stride 1 -> 1, stride 2 -> 0; the classifier's 7x7 stride-2 stem -> 2)
  g.pb = (ksize > stride ? ksize - stride : 0) / 2;
  g.wswap = wswap ? 1 : 0;
  g.alpha = alpha;
  g.act = act;
  return GS_OK;
}

bool tiled_ok(const ConvGeom& g) { return g.ksize == 3 && g.ci % 4 == 0 && g.co % 4 == 0; }

int grid_1d(size_t total, int block) {
  size_t b = (total + block - 1) / block;
  size_t cap = (size_t)gs_num_sms() * 32;
  return (int)(b < cap ? b : cap);
}

template <int S, int TN>
int launch_c(const float* x, const float* w, const float* bias, float* y, int n, int h, int wd, int kdim, int ndim,
             int oh, int ow, int pb, int w_is_kn, int flip, float alpha, int act, cudaStream_t st) {
  using T = CTile<S, TN>;
  auto kern = conv_c_tiled_kernel<S, TN>;
  static bool attr = false;
  if (!attr) {
    GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM));
    attr = true;
  }
  int th = gs_cdiv(oh, T::TH), tw = gs_cdiv(ow, T::TW);
  dim3 grid((unsigned)(n * th * tw), (unsigned)gs_cdiv(ndim, TN));
  kern<<<grid, T::NT, T::SMEM, st>>>(x, w, bias, y, n, h, wd, kdim, ndim, oh, ow, pb, w_is_kn, flip, alpha, act, th, tw);
  GS_CHECK_LAUNCH("conv_c_tiled");
  return GS_OK;
}

template <int TN>
int launch_t2(const float* in, const float* w, const float* bias, float* out, int n, int ih, int iw, int kdim,
              int ndim, int w_is_kn, float alpha, int act, cudaStream_t st) {
  using T = T2Tile<TN>;
  auto kern = conv_t2_tiled_kernel<TN>;
  static bool attr = false;
  if (!attr) {
    GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM));
    attr = true;
  }
  int th = gs_cdiv(ih, T::TH), tw = gs_cdiv(iw, T::TW);
  dim3 grid((unsigned)(n * th * tw), (unsigned)gs_cdiv(ndim, TN));
  kern<<<grid, T::NT, T::SMEM, st>>>(in, w, bias, out, n, ih, iw, kdim, ndim, w_is_kn, alpha, act, th, tw);
  GS_CHECK_LAUNCH("conv_t2_tiled");
  return GS_OK;
}

template <int S, int TC>
int launch_w(const float* big, const float* small, float* dw, int n, int h, int wd, int adim, int bdim, int oh,
             int ow, int pb, int out_ab, float alpha, cudaStream_t st) {
  using T = WTile<S, TC>;
  auto kern = conv_w_tiled_kernel<S, TC>;
  static bool attr = false;
  if (!attr) {
    GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM));
    attr = true;
  }
  int th = gs_cdiv(oh, T::TH), tw = gs_cdiv(ow, T::TW);
  int ntiles = n * th * tw;
  int ya = gs_cdiv(adim, TC), zb = gs_cdiv(bdim, TC);
  int per_sm = (T::SMEM > 110 * 1024) ? 1 : 2;
  int gx = (gs_num_sms() * per_sm * 2) / (ya * zb);
  if (gx < 1) gx = 1;
  if (gx > ntiles) gx = ntiles;
  dim3 grid((unsigned)gx, (unsigned)ya, (unsigned)zb);
  kern<<<grid, T::NT, T::SMEM, st>>>(big, small, dw, n, h, wd, adim, bdim, oh, ow, pb, out_ab, alpha, th, tw);
  GS_CHECK_LAUNCH("conv_w_tiled");
  return GS_OK;
}

// ------------------------------------------------------------------------------------------------
// 1x1 convolutions with 2 channels on one side (image side of the colour blocks)
bool pow2_le32(int v) { return v >= 1 && v <= 32 && (v & (v - 1)) == 0; }

// y[p][ndim] = act(alpha * x[p][kdim] . B + bias); returns 1 if handled
int try_conv1x1(const float* x, const float* w, const float* bias, float* y, long long npix, int kdim, int ndim,
                int w_is_kn, float alpha, int act, cudaStream_t st, bool* handled) {
  *handled = false;
  if (kdim == 2 && ndim % 4 == 0 && ndim <= 1024) {
    long long total = npix * (ndim / 4);
    int blocks = (int)((total + 256 * 4 - 1) / (256 * 4));
    if (blocks > gs_num_sms() * 16) blocks = gs_num_sms() * 16;
    if (blocks < 1) blocks = 1;
    conv1x1_expand_kernel<2><<<blocks, 256, (size_t)(3 * ndim) * sizeof(float), st>>>(x, w, bias, y, npix, ndim, w_is_kn, alpha, act);
    GS_CHECK_LAUNCH("conv1x1_expand");
    *handled = true;
  } else if (ndim == 2 && kdim % 4 == 0 && (pow2_le32(kdim / 4) || kdim % 128 == 0) && kdim <= 2048) {
    int lpp = kdim / 4 > 32 ? 32 : kdim / 4;
    long long warps = (npix + (32 / lpp) - 1) / (32 / lpp);
    long long blocks = (warps + 7) / 8 / 4 + 1;
    if (blocks > gs_num_sms() * 16) blocks = gs_num_sms() * 16;
    conv1x1_reduce_kernel<2><<<(int)blocks, 256, (size_t)(2 * kdim) * sizeof(float), st>>>(x, w, bias, y, npix, kdim, w_is_kn, alpha, act);
    GS_CHECK_LAUNCH("conv1x1_reduce");
    *handled = true;
  }
  return GS_OK;
}

// ------------------------------------------------------------------------------------------------
// tcgen05 path
// Pre-split (bf16 hi/lo) weights live in the CALLER's workspace (gs_context, common.cuh): an 8 MB scratch slot (weights
// used once: gradient-of-gradient operands) followed by a cache of PARAMETER weights, one slot per (parameter, layout).
// A parameter is used by several convolutions in the same orientation (D(real), D(fake), the penalty passes, both
// sub-steps): it is split at its first use and RE-SPLIT IN PLACE by gs_conv_weight_cache_refresh after every change of
// its value (optimiser update, load).  Slots never move, so kernels recorded in CUDA graphs stay valid, and the sub-steps
// contain no split kernels.  gs_conv_weight_cache_reset drops the table when the parameter buffers themselves move.
// Where the split copy of `w` in the layout (nt, kc tag, kn, flip) goes: the cached copy of a parameter (need_prep =
// false), a new cache entry, or the scratch slot.
unsigned char* prep_slot(gs_context* ctx, bool cacheable, const float* w, int kdim, int ndim, int nt, int kc, int kn, int flip,
                         size_t wbytes, bool* need_prep) {
  *need_prep = true;
  if (cacheable) {
    for (int i = 0; i < ctx->nprep; ++i) {
      const GsPrepKey& k = ctx->prep[i];
      if (k.w == w && k.kdim == kdim && k.ndim == ndim && k.nt == nt && k.kc == kc && k.kn == kn && k.flip == flip) {
        *need_prep = false;
        return ctx->ws + ctx->cache_off + k.off;
      }
    }
    if (ctx->nprep < 512 && ctx->used + wbytes <= ctx->cache) {
      ctx->prep[ctx->nprep++] = GsPrepKey{w, kdim, ndim, nt, kc, kn, flip, ctx->used};
      unsigned char* dst = ctx->ws + ctx->cache_off + ctx->used;
      ctx->used += (wbytes + 255) & ~(size_t)255;
      return dst;
    }
    if (!ctx->cache_full_warned) {
      ctx->cache_full_warned = 1;
      fprintf(stderr, "gansynth_b200: the split-weight cache of this context is full (%d entries, %zu of %zu bytes): "
                      "parameters are re-split at every use; pass a larger workspace\n", ctx->nprep, ctx->used, ctx->cache);
    }
  }
  return ctx->ws + ctx->scratch_off;
}

#ifdef GS_TC_PROF
// stage-profiling build only (tc_common.cuh): [role][wait0, wait1, wait2, total] cycle sums, filled by the kernels
unsigned long long* g_prof = nullptr;
unsigned long long* prof_table() {
  if (!g_prof) {
    cudaMalloc(&g_prof, 64 * 4 * sizeof(unsigned long long));
    cudaMemset(g_prof, 0, 64 * 4 * sizeof(unsigned long long));
  }
  return g_prof;
}
#define TC_SET_PROF(p) (p).prof = prof_table()
#else
#define TC_SET_PROF(p)
#endif

int tc_auto_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GS_CONV_TC");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v;
}

// shape gate of the tensor-core kernels: 3x3, contraction / output channels % 32, output channels <= 256,
// tile side a multiple of 16 rows (or 8 / 4 / 2 rows: images are interleaved) and of 8 columns
bool tc_ok(int form, int ksize, int kdim, int ndim, int th_dim, int tw_dim) {
  (void)form;
  if (ksize != 3 || kdim % 32 || ndim % 32 || ndim > 256) return false;
  const bool rows_ok = (th_dim % 16 == 0) || th_dim == 8 || th_dim == 4 || th_dim == 2;
  return rows_ok && tw_dim % 8 == 0;
}

template <int FORM, int KC>
int launch_tc_impl(const float* x, const float* w, const float* bias, float* y, int n, int h_in, int w_in, int h_out,
                   int w_out, int kdim, int ndim, int w_is_kn, int flip, float alpha, int act, bool cacheable,
                   cudaStream_t st, const Epi& epi, bool* fused) {
  using G = TcGeo<FORM>;
  const size_t wbytes = (size_t)9 * kdim * ndim * 2 * sizeof(__nv_bfloat16);
  GS_NEED_CONTEXT(ctx, "conv_tc");
  GS_CHECK_ARG(wbytes <= ctx->scratch, "conv_tc: weight of %zu bytes exceeds the scratch slot", wbytes);
  TcParams p;
  TC_SET_PROF(p);
  p.bias = bias; p.y = y;
  p.n_img = n; p.h_in = h_in; p.w_in = w_in; p.h_out = h_out; p.w_out = w_out; p.kdim = kdim; p.ndim = ndim;
  p.alpha = alpha; p.act = act;
  const int th_dim = (FORM == TC_T2) ? h_in : h_out, tw_dim = (FORM == TC_T2) ? w_in : w_out;
  if (th_dim % 16 == 0) { p.rows = 16; p.img = 1; p.tiles_h = th_dim / 16; }
  else { p.rows = th_dim; p.img = 16 / th_dim; p.tiles_h = 1; }
  p.tiles_w = tw_dim / 8;
  p.ntiles = ((n + p.img - 1) / p.img) * p.tiles_h * p.tiles_w;
  p.pix = G::pixels(p.rows, p.img);
  const int box_h = G::box_h(p.rows), box_w = G::box_w();
  p.rpix = box_h * p.img * box_w;
  for (int t = 0; t < 9; ++t) p.tap_off[t] = G::tap_off(t, p.img);
  // output-channel tile: all channels when there are enough pixel tiles to fill the GPU, else split so
  // that more SMs share the layer (each CTA re-reads the same halo, the weights are partitioned)
  int nt = ndim;
  if (G::NACC * nt > 512) nt = 512 / G::NACC;
  while (nt > 32 && nt % 64 == 0 && (long long)p.ntiles * (ndim / nt) < gs_num_sms()) nt /= 2;
  if (G::NACC == 1 && !getenv("GS_TC_OLD_NT")) {
    // Issue-time model (profiles/mma_timing_r1.txt, cycles per 16-channel K slice in cat mode: N=2nt MMA + N=nt MMA;
    // three N=256 MMAs at nt = 256): time ~ tiles per CTA x cycles per slice.  Halving the channel tile doubles the
    // CTAs but the thin MMAs are operand-fetch bound, so "split until every SM has a CTA" can overshoot (8x64 images:
    // 2 tiles x 107 cycles at nt = 32 against 1 x 135 at nt = 64).  Override only for a clear (>= 20 %) gain.
    auto cost = [](int t) { return t <= 32 ? 107.0 : t <= 64 ? 135.0 : t <= 128 ? 218.0 : 422.0; };
    auto model = [&](int t) {
      const int per_tile_ctas = std::max(1, gs_num_sms() / (ndim / t));
      return cost(t) * (double)((p.ntiles + per_tile_ctas - 1) / per_tile_ctas);
    };
    int best = nt;
    for (int t = 32; t <= ndim && t <= 256; t *= 2)
      if (ndim % t == 0 && model(t) < 0.8 * model(best)) best = t;
    nt = best;
  }
  size_t budget = 222 * 1024 - 1024 - 2 * 16384 - (FORM == TC_C2 ? 2048 : 0);   // alignment slack, output staging, parity table
  const size_t a_stage = (size_t)p.pix * KC * 4;
  const size_t raw = (((size_t)p.rpix * KC * 4) + 1023) & ~(size_t)1023;
  // MASK epilogue: a ring of 16 KB aux tiles comes out of the same budget (3 slots, 2 when the smallest ring
  // configuration would not fit beside them; else the mask runs as a separate pass)
  p.aux_k = 0;
  if (epi.mode == TC_EPI_MASK && !getenv("GS_TC_NO_EPI")) {
    const size_t minimal = a_stage + raw + 2 * (size_t)4 * KC * 32;
    for (int k = 3; k >= 2 && !p.aux_k; --k)
      if (minimal + (size_t)k * 16384 <= budget) p.aux_k = k;
    budget -= (size_t)p.aux_k * 16384;
  }
  p.raw_slot_bytes = (uint32_t)raw;
  p.sa = 2;
  p.ds = 2;
  // weights: resident when every (chunk, tap) block fits beside two raw slots; else streamed in the
  // largest tap group (9, 3, 1) that leaves room for two stages
  const int nchunks = kdim / KC;
  size_t b_tap = 0;
  for (;; nt /= 2) {
    b_tap = (size_t)4 * KC * nt;
    if (p.sa * a_stage + p.ds * raw + 2 * b_tap <= budget || nt <= 32) break;
  }
  // the widest staged tiles (stride-2 form on interleaved low-resolution images): single-buffered rings
  if (p.sa * a_stage + p.ds * raw + 2 * b_tap > budget) p.sa = 1;
  if (p.sa * a_stage + p.ds * raw + 2 * b_tap > budget) p.ds = 1;
  GS_CHECK_ARG(p.sa * a_stage + p.ds * raw + 2 * b_tap <= budget, "conv_tc: shared memory budget exceeded (ndim %d)", ndim);
  p.nt = nt;
  p.n_tiles = ndim / nt;
  size_t used = p.sa * a_stage + p.ds * raw;
  // split-K for layers that cannot give every SM a CTA (2- and 4-row 256-channel blocks: 16-64 CTAs, each walking
  // all 2304 contraction rows): grid.z CTAs take `cps` channel chunks each and add their partial tiles into the
  // pre-zeroed y with TMA reduce-add; bias rides on z = 0, the activation is a separate in-place pass afterwards
  p.ksplit = 1;
  {
    const long long ctas = (long long)std::min(p.ntiles, std::max(1, gs_num_sms() / p.n_tiles)) * p.n_tiles;
    if (2 * ctas <= gs_num_sms() && nchunks >= 2 && !getenv("GS_TC_NO_KSPLIT")) {
      const int want = (int)std::min<long long>(nchunks, gs_num_sms() / ctas);
      for (int ks = want; ks >= 2; --ks)
        if (nchunks % ks == 0) { p.ksplit = ks; break; }
    }
  }
  p.cps = nchunks / p.ksplit;
  // fused epilogue: MASK needs the whole contraction in one CTA, PNF also every output channel of a pixel
  p.epi = TC_EPI_PLAIN; p.eps = epi.eps; p.rvec = epi.rvec;
  if (epi.mode == TC_EPI_MASK && p.ksplit == 1 && p.aux_k > 0) p.epi = TC_EPI_MASK;
  if (epi.mode == TC_EPI_PNF && p.ksplit == 1 && nt == ndim) p.epi = TC_EPI_PNF;
  if (getenv("GS_TC_NO_EPI")) p.epi = TC_EPI_PLAIN;
  if (fused) *fused = (p.epi == epi.mode);
  p.b_resident = (p.cps <= TC_MAX_BSTAGES) && (used + (size_t)p.cps * 9 * b_tap <= budget) && !getenv("GS_TC_NO_RESIDENT");
  int tps = 1;
  if (p.b_resident) {
    tps = 9;
    p.sb = p.cps;
    used += (size_t)p.cps * 9 * b_tap;
  } else {
    if (used + 2 * 9 * b_tap <= budget) tps = 9;
    else if (used + 2 * 3 * b_tap <= budget) tps = 3;
    p.sb = 2;
    used += 2 * (size_t)tps * b_tap;
  }
  // spare room: a third raw slot (HBM latency) and a third operand stage (two issuing warps) first, then
  // a deeper weight ring, then a fourth raw slot / operand stage
  if (used + raw <= budget) { ++p.ds; used += raw; }
  if (used + a_stage <= budget) { ++p.sa; used += a_stage; }
  if (!p.b_resident)
    while (p.sb < 4 && used + (size_t)tps * b_tap <= budget) { ++p.sb; used += (size_t)tps * b_tap; }
  if (used + raw <= budget) { ++p.ds; used += raw; }
  if (used + a_stage <= budget) { ++p.sa; used += a_stage; }
  p.nbuf = (2 * G::NACC * nt <= 512) ? 2 : 1;
  p.cat = (4 * G::NACC * nt <= 512) && !getenv("GS_TC_NO_CAT");
  p.nw = (p.nbuf == 2 && getenv("GS_TC_TWO_MMA_WARPS") && p.epi != TC_EPI_MASK) ? 2 : 1;   // measured: a second issuing warp buys nothing
  if (p.nw == 2) {   // each issuing warp owns every other stage
    p.sa &= ~1;
    if (!p.b_resident) p.sb &= ~1;
  }
  int cols = 32;
  while (cols < p.nbuf * G::NACC * nt * (1 + p.cat)) cols <<= 1;
  p.tmem_cols = cols;
  {
    // pre-split weights: cached copy of a parameter, else split now (into the cache or the scratch slot)
    bool need_prep = true;
    unsigned char* dst = prep_slot(ctx, cacheable, w, kdim, ndim, nt, KC, w_is_kn, flip, wbytes, &need_prep);
    if (need_prep) {
      size_t total = (size_t)9 * kdim * ndim;
      int blocks = (int)((total + 255) / 256);
      if (blocks > gs_num_sms() * 8) blocks = gs_num_sms() * 8;
      conv_tc_prep_kernel<KC><<<blocks, 256, 0, st>>>(w, reinterpret_cast<__nv_bfloat16*>(dst), kdim, ndim, nt, w_is_kn, flip);
      GS_CHECK_LAUNCH("conv_tc_prep");
    }
    p.wprep = reinterpret_cast<const __nv_bfloat16*>(dst);
  }
  CUtensorMap tmx;
  TcOutMaps tmy, tmaux;
  {
    int rc = gs_make_act_tmap(&tmx, x, n, h_in, w_in, kdim, KC, box_w, p.img, box_h, KC * 4);
    if (rc) return rc;
    // output maps and, for the MASK epilogue, identical maps over the mask source
    for (int which = 0; which < 2; ++which) {
      TcOutMaps& tm = which ? tmaux : tmy;
      const float* base = which ? (p.epi == TC_EPI_MASK ? epi.aux : y) : y;
      if (FORM == TC_T2) {
        // four sub-pixel phases: strided views of y, one accumulator each
        for (int a = 0; a < 4; ++a) {
          rc = gs_make_act_tmap(&tm.m[a], base + ((size_t)(a >> 1) * w_out + (a & 1)) * ndim, n, h_out / 2, w_out / 2, ndim, 32, 8,
                                p.img, p.rows, 128, 2, 2);
          if (rc) return rc;
        }
      } else {
        rc = gs_make_act_tmap(&tm.m[0], base, n, h_out, w_out, ndim, 32, 8, p.img, p.rows, 128);
        if (rc) return rc;
        for (int a = 1; a < 4; ++a) tm.m[a] = tm.m[0];
      }
    }
  }
  const size_t smem = used + 1024 + 2 * 16384 + (size_t)p.aux_k * 16384 + (FORM == TC_C2 ? 2048 : 0);
  auto kern = conv_tc_kernel<FORM, KC, 9, 0>;
  if (p.cat) kern = tps == 9 ? conv_tc_kernel<FORM, KC, 9, 1> : tps == 3 ? conv_tc_kernel<FORM, KC, 3, 1> : conv_tc_kernel<FORM, KC, 1, 1>;
  else kern = tps == 9 ? conv_tc_kernel<FORM, KC, 9, 0> : tps == 3 ? conv_tc_kernel<FORM, KC, 3, 0> : conv_tc_kernel<FORM, KC, 1, 0>;
  static bool attr[6] = {false, false, false, false, false, false};
  const int ai = (tps == 9 ? 0 : tps == 3 ? 1 : 2) + 3 * p.cat;
  if (!attr[ai]) {
    GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
    attr[ai] = true;
  }
  int gx = gs_num_sms() / p.n_tiles;
  if (gx < 1) gx = 1;
  if (gx > p.ntiles) gx = p.ntiles;
  dim3 grid((unsigned)gx, (unsigned)p.n_tiles, (unsigned)p.ksplit);
  const size_t y_count = (size_t)n * h_out * w_out * ndim;
  if (p.ksplit > 1) GS_CUDA(cudaMemsetAsync(y, 0, y_count * sizeof(float), st));
  kern<<<grid, TC_THREADS, smem, st>>>(tmx, tmy, tmaux, p);
  GS_CHECK_LAUNCH("conv_tc");
  if (p.ksplit > 1 && act == 1) return gs_lrelu(y, y, (long long)y_count, st);
  return GS_OK;
}

// kw-stacked stride-1 kernel for thin layers (conv_tck.cuh): 32 or 64 output channels, all of them in one CTA,
// rows in multiples of 8, wide enough images that the 14-column tiles fill the GPU
bool tck_ok(int h, int w, int kdim, int ndim) {
  return (ndim == 32 || ndim == 64) && kdim % 32 == 0 && h % 8 == 0 && w >= 128 && !getenv("GS_TC_NO_KSTACK");
}

int launch_tck(const float* x, const float* w, const float* bias, float* y, int n, int h, int wd, int kdim, int ndim,
               int w_is_kn, int flip, float alpha, int act, bool cacheable, cudaStream_t st, const Epi& epi, bool* fused) {
  const size_t wbytes = (size_t)9 * kdim * ndim * 2 * sizeof(__nv_bfloat16);
  GS_NEED_CONTEXT(ctx, "conv_tck");
  GS_CHECK_ARG(wbytes <= ctx->scratch, "conv_tck: weight of %zu bytes exceeds the scratch slot", wbytes);
  TckParams p;
  TC_SET_PROF(p);
  p.bias = bias; p.n_img = n; p.h = h; p.w = wd; p.kdim = kdim; p.nt = ndim; p.alpha = alpha; p.act = act;
  p.tiles_h = h / 8; p.tiles_w = (wd + 13) / 14; p.ntiles = n * p.tiles_h * p.tiles_w;
  // "cat" (hi x [hi | lo] as one double-width MMA) halves the MMA count of a K slice but doubles the accumulator columns
  // the epilogue has to read back, and TMEM reads run at 64 B/clk per SM: 128 lanes x 6 nt columns = 1536 cycles per
  // tile at nt = 32, more than the MMAs of a 32-channel contraction take.  Cat only when the contraction is long
  // enough to hide that (GS_TCK_CAT=0/1 forces it for A/B runs).
  p.cat = (6 * ndim <= 256) && kdim > 32 && !getenv("GS_TC_NO_CAT");
  if (const char* e = getenv("GS_TCK_CAT")) p.cat = (e[0] == '1') && (6 * ndim <= 256);
  p.nbuf = std::min(4, 512 / ((p.cat ? 6 : 3) * ndim));
  p.epi = getenv("GS_TC_NO_EPI") ? TC_EPI_PLAIN : epi.mode;
  p.eps = epi.eps; p.rvec = epi.rvec;
  const int nchunks = kdim / 32;
  size_t budget = 222 * 1024 - 1024 - 2 * TCK_OUT;
  const size_t a_stage = 2 * 4 * TCK_PIX * 16, raw = TCK_RAW;
  const size_t b_row = (size_t)4 * 6 * ndim * 16;
  // MASK epilogue: ring of aux tiles (see launch_tc_impl); at least one output tile's worth of chunks
  p.aux_k = 0;
  if (p.epi == TC_EPI_MASK) {
    const size_t minimal = 2 * a_stage + 2 * raw + 3 * b_row;
    for (int k = 3; k >= 2 && !p.aux_k; --k)
      if (minimal + (size_t)k * TCK_OUT <= budget) p.aux_k = k;
    if (!p.aux_k) p.epi = TC_EPI_PLAIN;
    budget -= (size_t)p.aux_k * TCK_OUT;
  }
  if (fused) *fused = (p.epi == epi.mode);
  p.sa = 2; p.ds = 2;
  size_t used = p.sa * a_stage + p.ds * raw;
  int tps = 1;
  p.b_resident = (used + (size_t)nchunks * 3 * b_row <= budget) && nchunks <= TC_MAX_BSTAGES;
  if (p.b_resident) { tps = 3; p.sb = nchunks; used += (size_t)nchunks * 3 * b_row; }
  else if (used + 2 * 3 * b_row <= budget) { tps = 3; p.sb = 2; used += 2 * 3 * b_row; }
  else { tps = 1; p.sb = 3; used += 3 * b_row; }
  GS_CHECK_ARG(used <= budget, "conv_tck: shared memory budget exceeded (kdim %d ndim %d)", kdim, ndim);
  if (used + raw <= budget) { ++p.ds; used += raw; }
  if (used + a_stage <= budget) { ++p.sa; used += a_stage; }
  if (!p.b_resident)
    while (p.sb < 4 && used + (size_t)tps * b_row <= budget) { ++p.sb; used += (size_t)tps * b_row; }
  if (used + raw <= budget) { ++p.ds; used += raw; }
  if (used + a_stage <= budget) { ++p.sa; used += a_stage; }
  p.tmem_cols = 512;
  {
    const int layout_tag = 1032;      // distinguishes the kw-stacked layout from conv_tc's in the cache
    bool need_prep = true;
    unsigned char* dst = prep_slot(ctx, cacheable, w, kdim, ndim, ndim, layout_tag, w_is_kn, flip, wbytes, &need_prep);
    if (need_prep) {
      size_t total = (size_t)9 * kdim * ndim;
      int blocks = (int)((total + 255) / 256);
      if (blocks > gs_num_sms() * 8) blocks = gs_num_sms() * 8;
      conv_tck_prep_kernel<<<blocks, 256, 0, st>>>(w, reinterpret_cast<__nv_bfloat16*>(dst), kdim, ndim, w_is_kn, flip);
      GS_CHECK_LAUNCH("conv_tck_prep");
    }
    p.wprep = reinterpret_cast<const __nv_bfloat16*>(dst);
  }
  CUtensorMap tmx, tmy, tmaux;
  int rc = gs_make_act_tmap(&tmx, x, n, h, wd, kdim, 32, 16, 1, 10, 128);
  if (rc) return rc;
  rc = gs_make_act_tmap(&tmy, y, n, h, wd, ndim, 32, 14, 1, 8, 128);
  if (rc) return rc;
  rc = gs_make_act_tmap(&tmaux, p.epi == TC_EPI_MASK ? epi.aux : y, n, h, wd, ndim, 32, 14, 1, 8, 128);
  if (rc) return rc;
  using TckKernel = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TckParams);
  // [epilogue][cat][tps == 3]
  static const TckKernel kerns[3][2][2] = {
      {{conv_tck_kernel<1, 0, 0>, conv_tck_kernel<3, 0, 0>}, {conv_tck_kernel<1, 1, 0>, conv_tck_kernel<3, 1, 0>}},
      {{conv_tck_kernel<1, 0, 1>, conv_tck_kernel<3, 0, 1>}, {conv_tck_kernel<1, 1, 1>, conv_tck_kernel<3, 1, 1>}},
      {{conv_tck_kernel<1, 0, 2>, conv_tck_kernel<3, 0, 2>}, {conv_tck_kernel<1, 1, 2>, conv_tck_kernel<3, 1, 2>}}};
  TckKernel kern = kerns[p.epi][p.cat ? 1 : 0][tps == 3 ? 1 : 0];
  static bool attr[12] = {false};
  const int ai = (p.epi * 2 + (p.cat ? 1 : 0)) * 2 + (tps == 3 ? 1 : 0);
  if (!attr[ai]) {
    GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
    attr[ai] = true;
  }
  int gx = gs_num_sms();
  if (gx > p.ntiles) gx = p.ntiles;
  kern<<<gx, TCK_THREADS, used + 1024 + 2 * TCK_OUT + (size_t)p.aux_k * TCK_OUT, st>>>(tmx, tmy, tmaux, p);
  GS_CHECK_LAUNCH("conv_tck");
  return GS_OK;
}

// bf16x3 (two-term split).  A three-term split was measured and buys nothing: the tcgen05 fp32 accumulator
// truncates (tools/tc_precision.py: mean relative error -9.5e-6 at K = 2304 whatever the split), so the
// accumulation, not the operand split, bounds the accuracy at ~1e-5.
template <int FORM>
int launch_tc(const float* x, const float* w, const float* bias, float* y, int n, int h_in, int w_in, int h_out,
              int w_out, int kdim, int ndim, int w_is_kn, int flip, float alpha, int act, bool cacheable,
              cudaStream_t st, const Epi& epi = Epi(), bool* fused = nullptr) {
  if (FORM == TC_C1 && tck_ok(h_out, w_out, kdim, ndim))
    return launch_tck(x, w, bias, y, n, h_out, w_out, kdim, ndim, w_is_kn, flip, alpha, act, cacheable, st, epi, fused);
  // the stride-2 gather form stages ~4x the pixels of the others: 16-channel chunks keep the rings in budget
  if (FORM == TC_C2)
    return launch_tc_impl<FORM, 16>(x, w, bias, y, n, h_in, w_in, h_out, w_out, kdim, ndim, w_is_kn, flip, alpha, act, cacheable, st, epi, fused);
  return launch_tc_impl<FORM, 32>(x, w, bias, y, n, h_in, w_in, h_out, w_out, kdim, ndim, w_is_kn, flip, alpha, act, cacheable, st, epi, fused);
}

// filter-gradient form on the tensor cores
bool tcw_ok(int ksize, int adim, int bdim, int sh, int sw) {
  return ksize == 3 && adim % 32 == 0 && bdim % 32 == 0 && sw % 8 == 0 && sh % 2 == 0;
}

// largest channel block (<= 128, multiple of 32) that divides c
int tcw_block(int c) {
  for (int blk = 128; blk > 32; blk -= 32)
    if (c % blk == 0) return blk;
  return 32;
}

int launch_tcw(const float* big, const float* small, float* dw, int n, int bh, int bw, int adim, int bdim, int sh,
               int sw, int stride, int out_ab, float alpha, cudaStream_t st, float* dbias = nullptr, int bias_side = 0) {
  TcwParams p;
  TC_SET_PROF(p);
  p.dw = dw;
  p.dbias = dbias; p.bias_side = dbias ? bias_side : 0;
  p.n_img = n; p.bh = bh; p.bw = bw; p.sh = sh; p.sw = sw; p.adim = adim; p.bdim = bdim; p.stride = stride;
  p.out_ab = out_ab; p.alpha = alpha;
  p.nch = tcw_block(adim);
  p.nb = tcw_block(bdim);
  p.nstack = (stride == 1 && p.nb <= 32 && !getenv("GS_TCW_NO_NSTACK")) ? 1 : 0;   // measured: no gain at NB = 64
  p.cat = (p.nstack ? (6 * p.nb <= 256) : (3 * p.nb * 2 <= 512)) && !getenv("GS_TC_NO_CAT");
  p.njobs_n = bdim / p.nb;
  // M jobs: as many kh taps per 128-row accumulator as fit (16 chunks of 8 channels)
  const int qj = p.nch / 8, qb = p.nb / 8;
  const int nkh_job = qj <= 4 ? 3 : qj <= 8 ? 2 : 1;
  p.mjobs = 0;
  for (int c0 = 0; c0 < adim; c0 += p.nch)
    for (int k0 = 0; k0 < 3; k0 += nkh_job) {
      GS_CHECK_ARG(p.mjobs < 8, "conv_tcw: too many jobs (adim %d)", adim);
      p.kh0[p.mjobs] = k0;
      p.nkh[p.mjobs] = (3 - k0 < nkh_job) ? 3 - k0 : nkh_job;
      p.ch0[p.mjobs] = c0;
      p.map_id[p.mjobs] = (p.nkh[p.mjobs] == nkh_job) ? 0 : 1;
      ++p.mjobs;
    }
  p.pw = stride == 1 ? (p.nstack ? 8 : 10) : 18;
  p.bwraw = stride == 1 ? (p.nstack ? 8 : 10) : 17;
  const size_t budget = 222 * 1024 - 1024 - 8192;       // alignment slack; over-read slack of the 128-row M operand
  int tpr = 16;
  if (const char* e = getenv("GS_TCW_TPR")) tpr = atoi(e);      // experiment knob: largest tile height tried
  size_t stage = 0, raw = 0;
  for (;; tpr >>= 1) {
    GS_CHECK_ARG(tpr >= 2, "conv_tcw: no tile fits shared memory (adim %d bdim %d)", adim, bdim);
    if (sh % tpr) continue;
    p.hr_max = stride * (tpr - 1) + nkh_job;
    const size_t big_stage = (size_t)p.hr_max * qj * p.pw * 16;          // one split term
    const size_t small_stage = (size_t)tpr * (p.nstack ? 6 : 2) * qb * 128;
    p.big_lo_off = (uint32_t)big_stage;
    p.small_off = (uint32_t)(2 * big_stage);
    stage = 2 * big_stage + small_stage;
    p.raw_big_chunk = (uint32_t)((((size_t)p.hr_max * p.bwraw * 128) + 1023) & ~(size_t)1023);
    p.raw_small_chunk = (uint32_t)((((size_t)tpr * (p.nstack ? 10 : 8) * 128) + 1023) & ~(size_t)1023);
    raw = (size_t)(p.nch / 32) * p.raw_big_chunk + (size_t)(p.nb / 32) * p.raw_small_chunk;
    if (2 * stage + 2 * raw <= budget) break;
  }
  p.tpr = tpr;
  p.stage_bytes = (uint32_t)stage;
  p.raw_slot_bytes = (uint32_t)raw;
  p.stages = 2; p.ds = 2;
  size_t used = 2 * stage + 2 * raw;
  if (used + raw <= budget) { ++p.ds; used += raw; }
  if (used + stage <= budget) { ++p.stages; used += stage; }
  if (used + raw <= budget) { ++p.ds; used += raw; }
  p.tiles_h = sh / tpr; p.tiles_w = sw / 8; p.ntiles = n * p.tiles_h * p.tiles_w;
  {
    // fused bias gradient over the big operand: the kh jobs of a channel block load overlapping row ranges; hand
    // every tile row [0, stride * tpr) (relative to the tile's first image row) to exactly one of them
    const int pb = stride == 1 ? 1 : 0;
    p.bias_c0 = (stride == 1 && !p.nstack) ? 1 : 0;
    p.bias_c1 = stride == 1 ? p.bias_c0 + 8 : 16;
    int covered = 0, block = -1;
    for (int j = 0; j < p.mjobs; ++j) {
      if (p.ch0[j] != block) { block = p.ch0[j]; covered = 0; }
      const int first = p.kh0[j] - pb;                                   // tile-relative image row of raw row 0
      const int lo = std::max(first, covered), hi = std::min(first + stride * (tpr - 1) + p.nkh[j], stride * tpr);
      p.bias_r0[j] = lo - first;
      p.bias_r1[j] = hi > lo ? hi - first : lo - first;
      if (hi > lo) covered = hi;
    }
    for (int j = p.mjobs; j < 8; ++j) p.bias_r0[j] = p.bias_r1[j] = 0;
  }
  int cols = 32;
  while (cols < 3 * p.nb * (1 + p.cat)) cols <<= 1;       // same count stacked or not: 3 kw blocks x NB x (1 + cat)
  p.tmem_cols = cols;
  TcwMaps maps;
  {
    int rc = gs_make_act_tmap(&maps.big[0], big, n, bh, bw, adim, 32, p.bwraw, 1, stride * (tpr - 1) + nkh_job, 128);
    if (rc) return rc;
    const int nkh_last = 3 % nkh_job ? 3 % nkh_job : nkh_job;
    rc = gs_make_act_tmap(&maps.big[1], big, n, bh, bw, adim, 32, p.bwraw, 1, stride * (tpr - 1) + nkh_last, 128);
    if (rc) return rc;
    rc = gs_make_act_tmap(&maps.small, small, n, sh, sw, bdim, 32, p.nstack ? 10 : 8, 1, tpr, 128);
    if (rc) return rc;
  }
  const int njobs = p.mjobs * p.njobs_n;
  int px = gs_num_sms() / njobs;
  if (px < 1) px = 1;
  if (px > p.ntiles) px = p.ntiles;
  static bool attr = false;
  if (!attr) {
    GS_CUDA(cudaFuncSetAttribute(conv_tcw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
    attr = true;
  }
  dim3 grid((unsigned)px, (unsigned)njobs);
  conv_tcw_kernel<<<grid, TCW_THREADS, used + 1024 + 8192, st>>>(maps, p);
  GS_CHECK_LAUNCH("conv_tcw");
  return GS_OK;
}

}  // namespace

#ifdef GS_TC_PROF
// copies the table to `out` (64 x 4 values) and clears it; synchronises the device
extern "C" int gs_tc_prof_read(unsigned long long* out) {
  GS_CUDA(cudaDeviceSynchronize());
  GS_CUDA(cudaMemcpy(out, prof_table(), 64 * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  GS_CUDA(cudaMemset(prof_table(), 0, 64 * 4 * sizeof(unsigned long long)));
  return GS_OK;
}
#endif

// Re-splits, in place, every cached parameter weight that lies in [lo, hi) (both NULL: all of them): the caller has
// just changed those parameters (optimiser update, checkpoint load).  The cache slots keep their addresses, so kernels
// already recorded in CUDA graphs keep reading the right copies.  One launch per 48 entries.
extern "C" int gs_conv_weight_cache_refresh(const float* lo, const float* hi, void* stream) {
  gs_context* ctx = gs_bound_context();
  if (!ctx) return GS_OK;
  PrepBatch b;
  b.njobs = 0; b.nblocks = 0;
  auto flush = [&]() -> int {
    if (b.njobs == 0) return GS_OK;
    conv_prep_batch_kernel<<<b.nblocks, 256, 0, (cudaStream_t)stream>>>(b);
    GS_CHECK_LAUNCH("conv_prep_batch");
    b.njobs = 0; b.nblocks = 0;
    return GS_OK;
  };
  for (int i = 0; i < ctx->nprep; ++i) {
    const GsPrepKey& k = ctx->prep[i];
    if (lo != nullptr && (k.w < lo || k.w >= hi)) continue;
    PrepJob& j = b.jobs[b.njobs++];
    j.w = k.w;
    j.out = reinterpret_cast<__nv_bfloat16*>(ctx->ws + ctx->cache_off + k.off);
    j.kdim = k.kdim; j.ndim = k.ndim; j.nt = k.nt; j.kc = k.kc; j.kn = k.kn; j.flip = k.flip;
    j.block0 = b.nblocks;
    b.nblocks += (int)(((size_t)9 * k.kdim * k.ndim + PREP_ELEMS_PER_BLOCK - 1) / PREP_ELEMS_PER_BLOCK);
    if (b.njobs == 48) {
      int rc = flush();
      if (rc) return rc;
    }
  }
  return flush();
}

extern "C" int gs_conv_weight_cache_reset(void) {
  if (gs_context* ctx = gs_bound_context()) {
    ctx->nprep = 0;
    ctx->used = 0;
  }
  return GS_OK;
}

namespace {
// the un-fused form of an epilogue request, in place on the finished convolution output
int epi_fallback(const Epi& epi, float* y, long long pixels, int channels, cudaStream_t st) {
  if (epi.mode == TC_EPI_MASK) return gs_lrelu_mask_mul(y, epi.aux, y, pixels * channels, st);
  if (epi.mode == TC_EPI_PNF) return gs_pixel_norm_fwd(y, y, epi.rvec, pixels, channels, epi.eps, st);
  return GS_OK;
}

int check_epi(const Epi& epi, const float* bias, int act) {
  GS_CHECK_ARG(epi.mode >= 0 && epi.mode <= 2, "conv: unknown epilogue %d", epi.mode);
  GS_CHECK_ARG(epi.mode != TC_EPI_MASK || (epi.aux && !bias && act == 0), "conv: the mask epilogue takes a mask source, no bias, no activation");
  GS_CHECK_ARG(epi.mode != TC_EPI_PNF || (epi.rvec && act == 1), "conv: the pixel-norm epilogue needs rvec and act = 1 (leaky relu)");
  return GS_OK;
}

int conv_fwd_impl(const float* x, const float* w, const float* bias, float* y, int n, int h, int wd,
                  int ci, int co, int ksize, int stride, int wswap, float alpha, int act, int impl,
                  void* stream, const Epi& epi, bool* fused) {
  *fused = false;
  ConvGeom g;
  int rc = make_geom(g, n, h, wd, ci, co, ksize, stride, wswap, alpha, act);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool cacheable = (impl & GS_IMPL_PARAM_WEIGHT) != 0;   // w is a parameter: its bf16 split may be cached
  impl &= 0xff;
  bool tiled = tiled_ok(g);
  if (impl == 4) impl = tiled ? 2 : 1;   // "fp32 auto": never the tensor-core kernel
  GS_CHECK_ARG(!(impl == 2 && !tiled), "conv2d_fwd: tiled kernel needs ksize 3 and channels %% 4 == 0");
  {
    const int form = stride == 1 ? TC_C1 : TC_C2;
    const bool tcok = tc_ok(form, ksize, ci, co, g.oh, g.ow);
    GS_CHECK_ARG(!(impl == 3 && !tcok), "conv2d_fwd: tensor-core kernel does not cover this shape");
    if (impl == 3 || (impl == 0 && tcok && tc_auto_enabled())) {
      if (stride == 1) return launch_tc<TC_C1>(x, w, bias, y, n, h, wd, g.oh, g.ow, ci, co, !g.wswap, 0, alpha, act, cacheable, st, epi, fused);
      return launch_tc<TC_C2>(x, w, bias, y, n, h, wd, g.oh, g.ow, ci, co, !g.wswap, 0, alpha, act, cacheable, st, epi, fused);
    }
  }
  if (impl != 1 && ksize == 1 && stride == 1) {
    bool handled = false;
    rc = try_conv1x1(x, w, bias, y, (long long)n * h * wd, ci, co, !g.wswap, alpha, act, st, &handled);
    if (rc || handled) return rc;
  }
  if (impl == 1 || !tiled) {
    size_t total = (size_t)n * g.oh * g.ow * co;
    conv_c_naive_kernel<<<grid_1d(total, 256), 256, 0, st>>>(x, w, bias, y, g);
    GS_CHECK_LAUNCH("conv_c_naive");
    return GS_OK;
  }
  int kn = !g.wswap;
  if (stride == 1) {
    if (co <= 32) return launch_c<1, 32>(x, w, bias, y, n, h, wd, ci, co, g.oh, g.ow, g.pb, kn, 0, alpha, act, st);
    return launch_c<1, 64>(x, w, bias, y, n, h, wd, ci, co, g.oh, g.ow, g.pb, kn, 0, alpha, act, st);
  }
  if (co <= 32) return launch_c<2, 32>(x, w, bias, y, n, h, wd, ci, co, g.oh, g.ow, g.pb, kn, 0, alpha, act, st);
  return launch_c<2, 64>(x, w, bias, y, n, h, wd, ci, co, g.oh, g.ow, g.pb, kn, 0, alpha, act, st);
}

}  // namespace

extern "C" int gs_conv2d_fwd_ex(const float* x, const float* w, const float* bias, float* y, int n, int h, int wd, int ci,
                                int co, int ksize, int stride, int wswap, float alpha, int act, int epi_mode,
                                const float* aux, float* rvec, float eps, int impl, void* stream) {
  Epi epi;
  epi.mode = epi_mode; epi.aux = aux; epi.rvec = rvec; epi.eps = eps;
  int rc = check_epi(epi, bias, act);
  if (rc) return rc;
  bool fused = false;
  rc = conv_fwd_impl(x, w, bias, y, n, h, wd, ci, co, ksize, stride, wswap, alpha, act, impl, stream, epi, &fused);
  if (rc || fused) return rc;
  return epi_fallback(epi, y, (long long)n * (h / stride) * (wd / stride), co, (cudaStream_t)stream);
}

extern "C" int gs_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, int n, int h, int wd,
                             int ci, int co, int ksize, int stride, int wswap, float alpha, int act, int impl,
                             void* stream) {
  bool fused = false;
  return conv_fwd_impl(x, w, bias, y, n, h, wd, ci, co, ksize, stride, wswap, alpha, act, impl, stream, Epi(), &fused);
}

namespace {
int conv_dgrad_impl(const float* dy, const float* w, const float* bias, float* dx, int n, int h, int wd,
                    int ci, int co, int ksize, int stride, int wswap, float alpha, int act, int impl,
                    void* stream, const Epi& epi, bool* fused) {
  *fused = false;
  ConvGeom g;
  int rc = make_geom(g, n, h, wd, ci, co, ksize, stride, wswap, alpha, act);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool cacheable = (impl & GS_IMPL_PARAM_WEIGHT) != 0;
  impl &= 0xff;
  bool tiled = tiled_ok(g);
  if (impl == 4) impl = tiled ? 2 : 1;
  GS_CHECK_ARG(!(impl == 2 && !tiled), "conv2d_dgrad: tiled kernel needs ksize 3 and channels %% 4 == 0");
  {
    // contraction over co, output channels ci
    const int form = stride == 1 ? TC_C1 : TC_T2;
    const bool tcok = stride == 1 ? tc_ok(form, ksize, co, ci, h, wd) : tc_ok(form, ksize, co, ci, g.oh, g.ow);
    GS_CHECK_ARG(!(impl == 3 && !tcok), "conv2d_dgrad: tensor-core kernel does not cover this shape");
    if (impl == 3 || (impl == 0 && tcok && tc_auto_enabled())) {
      if (stride == 1) return launch_tc<TC_C1>(dy, w, bias, dx, n, h, wd, h, wd, co, ci, g.wswap, 1, alpha, act, cacheable, st, epi, fused);
      return launch_tc<TC_T2>(dy, w, bias, dx, n, g.oh, g.ow, h, wd, co, ci, g.wswap, 0, alpha, act, cacheable, st, epi, fused);
    }
  }
  if (impl != 1 && ksize == 1 && stride == 1) {
    bool handled = false;
    rc = try_conv1x1(dy, w, bias, dx, (long long)n * h * wd, co, ci, g.wswap, alpha, act, st, &handled);
    if (rc || handled) return rc;
  }
  if (impl == 1 || !tiled) {
    size_t total = (size_t)n * h * wd * ci;
    if (co >= 64 && total <= (size_t)gs_num_sms() * 256) {
      conv_t_naive_warp_kernel<<<grid_1d(total * 32, 256), 256, 0, st>>>(dy, w, bias, dx, g);
      GS_CHECK_LAUNCH("conv_t_naive_warp");
      return GS_OK;
    }
    conv_t_naive_kernel<<<grid_1d(total, 256), 256, 0, st>>>(dy, w, bias, dx, g);
    GS_CHECK_LAUNCH("conv_t_naive");
    return GS_OK;
  }
  // contraction over co, output channels ci: weight memory is [tap][K=co][N=ci] iff wswap
  int kn = g.wswap;
  if (stride == 1) {
    // dgrad of a stride-1 SAME conv == gather conv with the 180-degree rotated, transposed filter
    if (ci <= 32) return launch_c<1, 32>(dy, w, bias, dx, n, h, wd, co, ci, h, wd, 1, kn, 1, alpha, act, st);
    return launch_c<1, 64>(dy, w, bias, dx, n, h, wd, co, ci, h, wd, 1, kn, 1, alpha, act, st);
  }
  if (ci <= 32) return launch_t2<32>(dy, w, bias, dx, n, g.oh, g.ow, co, ci, kn, alpha, act, st);
  return launch_t2<64>(dy, w, bias, dx, n, g.oh, g.ow, co, ci, kn, alpha, act, st);
}

}  // namespace

extern "C" int gs_conv2d_dgrad_ex(const float* dy, const float* w, const float* bias, float* dx, int n, int h, int wd,
                                  int ci, int co, int ksize, int stride, int wswap, float alpha, int act, int epi_mode,
                                  const float* aux, float* rvec, float eps, int impl, void* stream) {
  Epi epi;
  epi.mode = epi_mode; epi.aux = aux; epi.rvec = rvec; epi.eps = eps;
  int rc = check_epi(epi, bias, act);
  if (rc) return rc;
  bool fused = false;
  rc = conv_dgrad_impl(dy, w, bias, dx, n, h, wd, ci, co, ksize, stride, wswap, alpha, act, impl, stream, epi, &fused);
  if (rc || fused) return rc;
  return epi_fallback(epi, dx, (long long)n * h * wd, ci, (cudaStream_t)stream);
}

extern "C" int gs_conv2d_dgrad(const float* dy, const float* w, const float* bias, float* dx, int n, int h, int wd,
                               int ci, int co, int ksize, int stride, int wswap, float alpha, int act, int impl,
                               void* stream) {
  bool fused = false;
  return conv_dgrad_impl(dy, w, bias, dx, n, h, wd, ci, co, ksize, stride, wswap, alpha, act, impl, stream, Epi(), &fused);
}

namespace {
int conv_wgrad_impl(const float* x, const float* dy, float* dw, int n, int h, int wd, int ci, int co, int ksize, int stride,
                    int wswap, float alpha, int impl, void* stream, float* dbias, int bias_of_x, bool* bias_done,
                    bool accumulate = false);
}
int gs_col_sum_acc(const float* v, float* out, long long rows, int c, void* stream);   // elementwise.cu: out += column sums

extern "C" int gs_conv2d_wgrad(const float* x, const float* dy, float* dw, int n, int h, int wd, int ci, int co,
                               int ksize, int stride, int wswap, float alpha, int impl, void* stream) {
  bool done = false;
  return conv_wgrad_impl(x, dy, dw, n, h, wd, ci, co, ksize, stride, wswap, alpha, impl, stream, nullptr, 0, &done);
}

// Filter gradient and, from the same pass over the operands, the bias gradient dbias[c] = sum over pixels of dy
// (bias_of_x = 0: conv2d layers, ops.py:240-244) or of x (bias_of_x = 1: the conv2d_transpose layers, whose
// pre-activation gradient is the HIGH-resolution operand, ops.py:272-277); dbias may be NULL.  accumulate = 0: dw and
// dbias are overwritten; 1: the results are ADDED to what they hold (several uses of one parameter summing into the
// caller's gradient buffer: tf.gradients' AddN without the extra passes).
extern "C" int gs_conv2d_wgrad_ex(const float* x, const float* dy, float* dw, float* dbias, int bias_of_x, int accumulate,
                                  int n, int h, int wd, int ci, int co, int ksize, int stride, int wswap, float alpha,
                                  int impl, void* stream) {
  const int cb = bias_of_x ? ci : co;
  if (dbias && !accumulate) GS_CUDA(cudaMemsetAsync(dbias, 0, (size_t)cb * sizeof(float), (cudaStream_t)stream));
  bool done = false;
  int rc = conv_wgrad_impl(x, dy, dw, n, h, wd, ci, co, ksize, stride, wswap, alpha, impl, stream, dbias, bias_of_x, &done,
                           accumulate != 0);
  if (rc || done || !dbias) return rc;
  const long long rows = bias_of_x ? (long long)n * h * wd : (long long)n * (h / stride) * (wd / stride);
  return gs_col_sum_acc(bias_of_x ? x : dy, dbias, rows, cb, stream);      // adds to the (zeroed or live) dbias
}

namespace {
int conv_wgrad_impl(const float* x, const float* dy, float* dw, int n, int h, int wd, int ci, int co, int ksize, int stride,
                    int wswap, float alpha, int impl, void* stream, float* dbias, int bias_of_x, bool* bias_done,
                    bool accumulate) {
  // every kernel below ADDS into dw with atomics: overwrite = zero first, accumulate = don't
  *bias_done = false;
  ConvGeom g;
  int rc = make_geom(g, n, h, wd, ci, co, ksize, stride, wswap, alpha, 0);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  bool tiled = tiled_ok(g);
  if (impl == 4) impl = tiled ? 2 : 1;
  GS_CHECK_ARG(!(impl == 2 && !tiled), "conv2d_wgrad: tiled kernel needs ksize 3 and channels %% 4 == 0");
  {
    const bool tcok = tcw_ok(ksize, ci, co, g.oh, g.ow);
    GS_CHECK_ARG(!(impl == 3 && !tcok), "conv2d_wgrad: tensor-core kernel does not cover this shape");
    if (impl == 3 || (impl == 0 && tcok && tc_auto_enabled())) {
      if (!accumulate) GS_CUDA(cudaMemsetAsync(dw, 0, (size_t)ksize * ksize * ci * co * sizeof(float), (cudaStream_t)stream));
      *bias_done = dbias != nullptr;
      return launch_tcw(x, dy, dw, n, h, wd, ci, co, g.oh, g.ow, stride, !g.wswap, alpha, (cudaStream_t)stream, dbias,
                        bias_of_x ? 2 : 1);
    }
  }
  size_t nel = (size_t)ksize * ksize * ci * co;
  if (!accumulate) GS_CUDA(cudaMemsetAsync(dw, 0, nel * sizeof(float), st));
  if (impl != 1 && ksize == 1 && stride == 1 && (ci == 2 || co == 2)) {
    const int W = (ci == 2) ? co : ci;
    if (W != 2 && W % 4 == 0 && W <= 256 && 256 % W == 0) {
      const long long npix = (long long)n * h * wd;
      // wide-channel index c, narrow index j -> element offset in dw ([ci][co], or [co][ci] when wswap)
      int sc, sj;
      if (ci == 2) { sc = g.wswap ? ci : 1; sj = g.wswap ? 1 : co; }
      else { sc = g.wswap ? 1 : co; sj = g.wswap ? ci : 1; }
      long long ppb = (npix + (long long)gs_num_sms() * 8 - 1) / ((long long)gs_num_sms() * 8);
      if (ppb < 64) ppb = 64;
      int blocks = (int)((npix + ppb - 1) / ppb);
      if (ci == 2) conv1x1_w_kernel<2><<<blocks, 256, 0, st>>>(dy, x, dw, npix, W, sc, sj, alpha, ppb);
      else conv1x1_w_kernel<2><<<blocks, 256, 0, st>>>(x, dy, dw, npix, W, sc, sj, alpha, ppb);
      GS_CHECK_LAUNCH("conv1x1_w");
      return GS_OK;
    }
  }
  if (impl == 1 || !tiled) {
    long long npix = (long long)n * g.oh * g.ow;
    int chunk = 512;
    {   // few pixels: shorter chunks so that the grid still fills the GPU
      long long bx = (long long)(nel + 127) / 128;
      long long want = ((long long)gs_num_sms() * 4 + bx - 1) / bx;
      long long c2 = (npix + want - 1) / want;
      if (c2 < 16) c2 = 16;
      if (c2 < chunk) chunk = (int)c2;
    }
    dim3 grid((unsigned)gs_cdiv((long long)nel, 128), (unsigned)gs_cdiv(npix, chunk));
    conv_w_naive_kernel<<<grid, 128, 0, st>>>(x, dy, dw, g, chunk);
    GS_CHECK_LAUNCH("conv_w_naive");
    return GS_OK;
  }
  int out_ab = !g.wswap;
  bool small = (ci <= 32 && co <= 32);
  if (stride == 1) {
    if (small) return launch_w<1, 32>(x, dy, dw, n, h, wd, ci, co, g.oh, g.ow, g.pb, out_ab, alpha, st);
    return launch_w<1, 64>(x, dy, dw, n, h, wd, ci, co, g.oh, g.ow, g.pb, out_ab, alpha, st);
  }
  if (small) return launch_w<2, 32>(x, dy, dw, n, h, wd, ci, co, g.oh, g.ow, g.pb, out_ab, alpha, st);
  return launch_w<2, 64>(x, dy, dw, n, h, wd, ci, co, g.oh, g.ow, g.pb, out_ab, alpha, st);
}
}  // namespace

// conv2d_transpose (ops.py:250-280): value [n,h,w,cin], variable [k,k,cin,filters], output
// [n,h*s,w*s,filters].  It is the dgrad form with the channel roles swapped.
extern "C" int gs_conv2d_transpose_fwd(const float* x, const float* var, const float* bias, float* y, int n, int h,
                                       int wd, int cin, int filters, int ksize, int stride, float alpha, int act,
                                       int impl, void* stream) {
  return gs_conv2d_dgrad(x, var, bias, y, n, h * stride, wd * stride, filters, cin, ksize, stride, 1, alpha, act, impl,
                         stream);
}
extern "C" int gs_conv2d_transpose_dgrad(const float* dy, const float* var, float* dx, int n, int h, int wd, int cin,
                                         int filters, int ksize, int stride, float alpha, int impl, void* stream) {
  return gs_conv2d_fwd(dy, var, nullptr, dx, n, h * stride, wd * stride, filters, cin, ksize, stride, 1, alpha, 0, impl,
                       stream);
}
extern "C" int gs_conv2d_transpose_wgrad(const float* x, const float* dy, float* dvar, int n, int h, int wd, int cin,
                                         int filters, int ksize, int stride, float alpha, int impl, void* stream) {
  return gs_conv2d_wgrad(dy, x, dvar, n, h * stride, wd * stride, filters, cin, ksize, stride, 1, alpha, impl, stream);
}
