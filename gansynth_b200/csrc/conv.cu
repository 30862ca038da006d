// C-ABI entry points for the NHWC convolution family (see include/gansynth_b200.h).
// Replaces the tf.nn.conv2d / tf.nn.conv2d_transpose call sites of the reference
// (ops.py:237, ops.py:269) and their TF-generated gradients (models.py:47,60,81-89).
#include "conv_tiled.cuh"
#include "gansynth_b200.h"

namespace {

int make_geom(ConvGeom& g, int n, int h, int w, int ci, int co, int ksize, int stride, int wswap, float alpha,
              int act) {
  GS_CHECK_ARG(n > 0 && h > 0 && w > 0 && ci > 0 && co > 0, "conv: non-positive dimension");
  GS_CHECK_ARG(ksize == 1 || ksize == 3, "conv: ksize must be 1 or 3 (got %d)", ksize);
  GS_CHECK_ARG(stride == 1 || stride == 2, "conv: stride must be 1 or 2 (got %d)", stride);
  GS_CHECK_ARG(h % stride == 0 && w % stride == 0, "conv: spatial size %dx%d not divisible by stride %d", h, w, stride);
  GS_CHECK_ARG(act == 0 || act == 1, "conv: act must be 0 (none) or 1 (leaky relu)");
  g.n = n; g.h = h; g.w = w; g.ci = ci; g.co = co;
  g.oh = h / stride; g.ow = w / stride;
  g.ksize = ksize; g.stride = stride;
  g.pb = (ksize == 3 && stride == 1) ? 1 : 0;  // TF SAME, even sizes
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

}  // namespace

extern "C" int gs_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, int n, int h, int wd,
                             int ci, int co, int ksize, int stride, int wswap, float alpha, int act, int impl,
                             void* stream) {
  ConvGeom g;
  int rc = make_geom(g, n, h, wd, ci, co, ksize, stride, wswap, alpha, act);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  bool tiled = tiled_ok(g);
  GS_CHECK_ARG(!(impl == 2 && !tiled), "conv2d_fwd: tiled kernel needs ksize 3 and channels %% 4 == 0");
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

extern "C" int gs_conv2d_dgrad(const float* dy, const float* w, const float* bias, float* dx, int n, int h, int wd,
                               int ci, int co, int ksize, int stride, int wswap, float alpha, int act, int impl,
                               void* stream) {
  ConvGeom g;
  int rc = make_geom(g, n, h, wd, ci, co, ksize, stride, wswap, alpha, act);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  bool tiled = tiled_ok(g);
  GS_CHECK_ARG(!(impl == 2 && !tiled), "conv2d_dgrad: tiled kernel needs ksize 3 and channels %% 4 == 0");
  if (impl == 1 || !tiled) {
    size_t total = (size_t)n * h * wd * ci;
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

extern "C" int gs_conv2d_wgrad(const float* x, const float* dy, float* dw, int n, int h, int wd, int ci, int co,
                               int ksize, int stride, int wswap, float alpha, int impl, void* stream) {
  ConvGeom g;
  int rc = make_geom(g, n, h, wd, ci, co, ksize, stride, wswap, alpha, 0);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  bool tiled = tiled_ok(g);
  GS_CHECK_ARG(!(impl == 2 && !tiled), "conv2d_wgrad: tiled kernel needs ksize 3 and channels %% 4 == 0");
  size_t nel = (size_t)ksize * ksize * ci * co;
  GS_CUDA(cudaMemsetAsync(dw, 0, nel * sizeof(float), st));
  if (impl == 1 || !tiled) {
    long long npix = (long long)n * g.oh * g.ow;
    int chunk = 512;
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
