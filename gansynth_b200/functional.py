"""Autograd layer over the CUDA kernels.

Every op is a torch.autograd.Function whose backward is built from OTHER Functions of this module, so
the R1 penalty (models.py:46-49 of the reference: gradient of D w.r.t. real images, differentiated
again w.r.t. D's weights) and the mode-seeking term (models.py:59-62: gradient of G w.r.t. latents,
differentiated w.r.t. G's weights) run entirely on the hand-written kernels:

* the convolution trio (gather C, transposed T, weight-gradient W) is bilinear and closed under
  differentiation, likewise the dense trio;
* leaky-relu has zero curvature, so only its mask (recovered from the sign of the output) is reused;
* pixel-norm, tanh and minibatch-stddev carry hand-derived second-derivative kernels.

PyTorch supplies the graph bookkeeping, device memory and streams only.  `K` is the kernel backend;
it is a CudaBackend in the product.  (tests/ swap in a torch-CPU emulation of the same primitive API to
check this module's graph logic against the oracle without a GPU.)
"""
import contextlib
import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .kernels import CudaBackend

K = CudaBackend()

ACT_NONE, ACT_LRELU = 0, 1


def set_backend(backend):
    global K
    K = backend


_SKIP_WGRAD = False
# GS_NO_FUSED_EW=1 keeps the un-fused elementwise chain (MaskMul, ColSum, PixelNorm) for A/B runs
FUSED_EW = os.environ.get("GS_NO_FUSED_EW", "0") != "1"
FUSED_MC = FUSED_EW and os.environ.get("GS_NO_FUSED_MC", "0") != "1"
FUSED_WB = FUSED_EW and os.environ.get("GS_NO_FUSED_WB", "0") != "1"      # bias gradient inside the filter-gradient kernel


@contextlib.contextmanager
def skip_weight_grads():
    """Inside this context the layer Functions do not compute parameter gradients in their backward.
    Used around the first-order passes of the two penalties (gradient w.r.t. images / latents only):
    the parameter gradients of that pass are never consumed."""
    global _SKIP_WGRAD
    prev, _SKIP_WGRAD = _SKIP_WGRAD, True
    try:
        yield
    finally:
        _SKIP_WGRAD = prev


# ----------------------------------------------------------------------------- convolution trio
class ConvC(Function):
    """y = alpha * gather_conv(x, w)  (forward form of tf.nn.conv2d, ops.py:237)."""

    @staticmethod
    def forward(ctx, x, w, ksize, stride, wswap, alpha):
        ctx.cfg = (ksize, stride, wswap, alpha)
        ctx.save_for_backward(x, w)
        return K.conv_c(x, w, None, ksize, stride, wswap, alpha, ACT_NONE)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        cfg = ctx.cfg
        dx = ConvT.apply(dy, w, *cfg) if ctx.needs_input_grad[0] else None
        dw = ConvW.apply(x, dy, *cfg) if ctx.needs_input_grad[1] else None
        return dx, dw, None, None, None, None


class ConvT(Function):
    """dx = alpha * transposed_conv(dy, w)  (input-gradient form; forward of tf.nn.conv2d_transpose)."""

    @staticmethod
    def forward(ctx, dy, w, ksize, stride, wswap, alpha):
        ctx.cfg = (ksize, stride, wswap, alpha)
        ctx.save_for_backward(dy, w)
        return K.conv_t(dy, w, None, ksize, stride, wswap, alpha, ACT_NONE)

    @staticmethod
    def backward(ctx, g):
        dy, w = ctx.saved_tensors
        cfg = ctx.cfg
        gdy = ConvC.apply(g, w, *cfg) if ctx.needs_input_grad[0] else None
        gw = ConvW.apply(g, dy, *cfg) if ctx.needs_input_grad[1] else None
        return gdy, gw, None, None, None, None


class ConvW(Function):
    """dw = alpha * sum_pixels x (x) dy  (filter-gradient form)."""

    @staticmethod
    def forward(ctx, x, dy, ksize, stride, wswap, alpha):
        ctx.cfg = (ksize, stride, wswap, alpha)
        ctx.save_for_backward(x, dy)
        return K.conv_w(x, dy, ksize, stride, wswap, alpha)

    @staticmethod
    def backward(ctx, gw):
        x, dy = ctx.saved_tensors
        cfg = ctx.cfg
        gx = ConvT.apply(dy, gw, *cfg) if ctx.needs_input_grad[0] else None
        gdy = ConvC.apply(x, gw, *cfg) if ctx.needs_input_grad[1] else None
        return gx, gdy, None, None, None, None


class ConvWB(Function):
    """(dw, db) = (ConvW(x, dy), column sum of the layer's pre-activation gradient) from ONE pass over the operands
    (gs_conv2d_wgrad_ex).  `bias_of`: 'dy' for conv2d layers, 'x' for conv2d_transpose layers (there the pre-activation
    gradient is the high-resolution operand `x` of the filter-gradient form)."""

    @staticmethod
    def forward(ctx, x, dy, ksize, stride, wswap, alpha, bias_of):
        ctx.cfg = (ksize, stride, wswap, alpha)
        ctx.bias_of = bias_of
        ctx.lead = tuple((x if bias_of == "x" else dy).shape[:-1])
        ctx.save_for_backward(x, dy)
        return K.conv_w(x, dy, ksize, stride, wswap, alpha, bias_of=bias_of)

    @staticmethod
    def backward(ctx, gw, gb):
        x, dy = ctx.saved_tensors
        cfg = ctx.cfg
        gx = ConvT.apply(dy, gw, *cfg) if ctx.needs_input_grad[0] else None
        gdy = ConvC.apply(x, gw, *cfg) if ctx.needs_input_grad[1] else None
        if ctx.bias_of == "x" and ctx.needs_input_grad[0]:
            gx = gx + RowBroadcast.apply(gb, ctx.lead)
        if ctx.bias_of == "dy" and ctx.needs_input_grad[1]:
            gdy = gdy + RowBroadcast.apply(gb, ctx.lead)
        return gx, gdy, None, None, None, None, None


# Gradient sinks: data_ptr of a parameter (a view of the packed parameter buffer) -> the view of the packed GRADIENT
# buffer laid out like it.  While a sink map is active (models.GANSynth._backward) and the backward pass is a plain
# first-order one, the convolution layers ADD their filter / bias gradients straight into the sinks
# (gs_conv2d_wgrad_ex, accumulate = 1) and hand autograd None: the three uses of a discriminator weight then cost no
# zero-fill, no AddN pass and no copy into the flat buffer.
_GRAD_SINKS = None


@contextlib.contextmanager
def grad_sinks(mapping):
    global _GRAD_SINKS
    prev, _GRAD_SINKS = _GRAD_SINKS, (mapping if FUSED_WB else None)
    try:
        yield
    finally:
        _GRAD_SINKS = prev


def _sink(key):
    """The gradient sink of the parameter whose data_ptr is `key`, if sinks are active in a first-order backward."""
    if _GRAD_SINKS is None or key is None or torch.is_grad_enabled():
        return None
    return _GRAD_SINKS.get(key)


def _wgrad_and_bias(form, x, dz, cfg, want_dw, want_db, w_key=None, b_key=None):
    """Parameter gradients of a conv layer in gather ('c') or transposed ('t') form from its input x and its
    pre-activation gradient dz -> (dw, db); one kernel when both are wanted.  With active gradient sinks for the
    parameters (`w_key` / `b_key`: their data_ptr) the results are accumulated there and (None, None) is returned."""
    dw = db = None
    w_sink = _sink(w_key) if want_dw else None
    if w_sink is not None:
        b_sink = _sink(b_key) if want_db else None
        if not want_db or b_sink is not None:
            if form == "c":
                K.conv_w(x, dz, *cfg, bias_of="dy" if want_db else None, out=(w_sink, b_sink))
            else:
                K.conv_w(dz, x, *cfg, bias_of="x" if want_db else None, out=(w_sink, b_sink))
            return None, None
    if want_dw and want_db and FUSED_WB:
        dw, db = ConvWB.apply(x, dz, *cfg, "dy") if form == "c" else ConvWB.apply(dz, x, *cfg, "x")
        return dw, db
    if want_dw:
        dw = ConvW.apply(x, dz, *cfg) if form == "c" else ConvW.apply(dz, x, *cfg)
    if want_db:
        db = ColSum.apply(dz)
    return dw, db


class PreMasked(object):
    """Wrapper protocol (D side).  `t` is the output y = leaky_relu(z) of a ConvLayer built with premasked=True: whoever
    consumes it returns the gradient w.r.t. the layer's PRE-activation z, i.e. lrelu'(y) * (gradient w.r.t. y) -- a
    convolution does that in its epilogue (ConvMasked), anything else goes through `plain()`."""
    __slots__ = ("t",)

    def __init__(self, t):
        self.t = t


class PnOut(object):
    """Wrapper protocol (G side).  `t` is the output y = pixel_norm(leaky_relu(z)) of a ConvPnLayer, `r` its per-pixel
    1 / sqrt(mean(a^2) + eps): whoever consumes it returns the gradient w.r.t. the layer's pre-activation z
    (PnBwdMaskY of the gradient w.r.t. y)."""
    __slots__ = ("t", "r")

    def __init__(self, t, r):
        self.t, self.r = t, r


def plain(x):
    """A protocol wrapper as an ordinary tensor for consumers that do not speak the protocol: identity forward, the
    un-fused mask / pixel-norm backward kernel on the way back."""
    if isinstance(x, PreMasked):
        return ApplyMask.apply(x.t)
    if isinstance(x, PnOut):
        return PnAdapter.apply(x.t, x.r)
    return x


class ConvLayer(Function):
    """Fused layer forward: act(alpha * conv(x, w) + bias), conv in gather (`form`='c') or transposed
    ('t') form.  The backward un-fuses into MaskMul(+ColSum) / the trio so it stays differentiable.
    `premasked`: the consumer of y (PixelNormOfLayer, or a consumer of PreMasked(y)) hands back a gradient that
    already carries the leaky-relu mask of this layer, so the backward must not apply it again.
    `x_premasked`: x is the output of such a premasked layer: the input gradient is returned multiplied by
    lrelu'(x), in the epilogue of the convolution that computes it (ConvMasked)."""

    @staticmethod
    def forward(ctx, x, w, bias, form, ksize, stride, wswap, alpha, act, premasked=False, x_premasked=False):
        fn = K.conv_c if form == "c" else K.conv_t
        y = fn(x, w, bias, ksize, stride, wswap, alpha, act, precise=True)
        ctx.cfg = (ksize, stride, wswap, alpha)
        ctx.form, ctx.act, ctx.has_bias, ctx.premasked = form, act, bias is not None, premasked
        ctx.x_premasked = x_premasked
        ctx.w_key, ctx.b_key = w.data_ptr(), (bias.data_ptr() if bias is not None else None)
        ctx.save_for_backward(x, w, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        cfg = ctx.cfg
        want_db = ctx.has_bias and ctx.needs_input_grad[2] and not _SKIP_WGRAD
        db = None
        if ctx.act == ACT_LRELU and not ctx.premasked:
            if want_db and FUSED_MC:
                dz, db = MaskMulColSum.apply(dy, y)      # mask and bias gradient in one pass
            else:
                dz = MaskMul.apply(dy, y)
        else:
            dz = dy
        dx = dw = None
        if ctx.needs_input_grad[0]:
            if ctx.x_premasked:
                dx = ConvMasked.apply(dz, w, x, "t" if ctx.form == "c" else "c", *cfg)
            else:
                dx = (ConvT if ctx.form == "c" else ConvC).apply(dz, w, *cfg)
        want_dw = ctx.needs_input_grad[1] and not _SKIP_WGRAD
        if want_dw or (want_db and db is None):
            dw, db2 = _wgrad_and_bias(ctx.form, x, dz, cfg, want_dw, want_db and db is None, ctx.w_key, ctx.b_key)
            db = db if db is not None else db2
        return dx, dw, db, None, None, None, None, None, None, None, None


class ConvMasked(Function):
    """lrelu'(src) * alpha * conv(v, w), conv in gather ('c') or transposed ('t') form, the mask applied in the
    convolution's epilogue (GS_EPI_MASK).  Linear in v and in w; `src` only contributes its sign."""

    @staticmethod
    def forward(ctx, v, w, src, form, ksize, stride, wswap, alpha):
        ctx.cfg = (ksize, stride, wswap, alpha)
        ctx.form = form
        ctx.save_for_backward(v, w, src)
        fn = K.conv_c if form == "c" else K.conv_t
        return fn(v, w, None, ksize, stride, wswap, alpha, ACT_NONE, mask_src=src)

    @staticmethod
    def backward(ctx, u):
        v, w, src = ctx.saved_tensors
        cfg = ctx.cfg
        mu = MaskMul.apply(u, src)
        gv = gw = None
        if ctx.needs_input_grad[0]:
            gv = (ConvT if ctx.form == "c" else ConvC).apply(mu, w, *cfg)
        if ctx.needs_input_grad[1]:
            gw = ConvW.apply(v, mu, *cfg) if ctx.form == "c" else ConvW.apply(mu, v, *cfg)
        return gv, gw, None, None, None, None, None, None


class ApplyMask(Function):
    """Identity on the output y of a premasked leaky-relu layer; the backward applies that layer's mask (the un-fused
    route of the PreMasked protocol)."""

    @staticmethod
    def forward(ctx, y):
        ctx.save_for_backward(y)
        return y.view_as(y)

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        return MaskMul.apply(g, y)


class ConvPnLayer(Function):
    """pixel_normalization(leaky_relu(alpha * conv(x, w) + bias)) as ONE kernel (networks.py:57-68, 82-93; the
    normalisation runs in the convolution's epilogue) -> (y, r).  The layer keeps y and r only.
    Protocol (PnOut): the incoming gradient is w.r.t. this layer's PRE-activation; `xr` is the r of the PnOut that x
    came from (None for an ordinary tensor), in which case the returned input gradient is w.r.t. THAT layer's
    pre-activation."""

    @staticmethod
    def forward(ctx, x, xr, w, bias, form, ksize, stride, wswap, alpha, eps):
        y, r = K.conv_pn(x, w, bias, form, ksize, stride, wswap, alpha, eps)
        ctx.cfg = (ksize, stride, wswap, alpha)
        ctx.form, ctx.has_bias = form, bias is not None
        ctx.w_key, ctx.b_key = w.data_ptr(), (bias.data_ptr() if bias is not None else None)
        ctx.save_for_backward(x, xr, w)
        ctx.mark_non_differentiable(r)
        return y, r

    @staticmethod
    def backward(ctx, dz, _gr):
        x, xr, w = ctx.saved_tensors
        cfg = ctx.cfg
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = (ConvT if ctx.form == "c" else ConvC).apply(dz, w, *cfg)
            if xr is not None:
                dx = PnBwdMaskY.apply(x, xr, dx)
        dw, db = _wgrad_and_bias(ctx.form, x, dz, cfg, ctx.needs_input_grad[2] and not _SKIP_WGRAD,
                                 ctx.has_bias and ctx.needs_input_grad[3] and not _SKIP_WGRAD, ctx.w_key, ctx.b_key)
        return dx, None, dw, db, None, None, None, None, None, None


class PnAdapter(Function):
    """Identity on the y of a PnOut; the backward turns the gradient w.r.t. y into the gradient w.r.t. the producing
    layer's pre-activation (the un-fused route of the PnOut protocol)."""

    @staticmethod
    def forward(ctx, y, r):
        ctx.save_for_backward(y, r)
        return y.view_as(y)

    @staticmethod
    def backward(ctx, g):
        y, r = ctx.saved_tensors
        return PnBwdMaskY.apply(y, r, g), None


class PnBwdMaskY(Function):
    """PnBwdMask for a layer that kept (y = a * r, r): dz = lrelu'(y) * pixel_norm_backward(y / r, r, dy).  Second
    order exactly as PnBwdMask (the gradient w.r.t. y's slot is the one w.r.t. the pre-activation)."""

    @staticmethod
    def forward(ctx, y, r, dy):
        ctx.save_for_backward(y, r, dy)
        return K.pn_bwd_mask_y(y, r, dy, False)[0]

    @staticmethod
    @once_differentiable
    def backward(ctx, u):
        y, r, dy = ctx.saved_tensors
        ga, gdy = K.pn_bwd_mask_second_y(y, r, dy, u)
        return ga, None, gdy


# ----------------------------------------------------------------------------- activations / bias
class MaskMul(Function):
    """v * lrelu'(.) with the mask taken from the sign of the layer OUTPUT y (same sign as the input)."""

    @staticmethod
    def forward(ctx, v, y):
        ctx.save_for_backward(y)
        return K.mask_mul(v, y)

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        return MaskMul.apply(g, y), None


class MaskMulColSum(Function):
    """(v * lrelu'(y), bias gradient = column sums of that product) in one pass.  The bias gradient only
    feeds the optimiser: it is never differentiated."""

    @staticmethod
    def forward(ctx, v, y):
        ctx.save_for_backward(y)
        out, cs = K.mask_mul_colsum(v, y)
        ctx.mark_non_differentiable(cs)
        return out, cs

    @staticmethod
    def backward(ctx, g, _gcs):
        (y,) = ctx.saved_tensors
        return MaskMul.apply(g, y), None


class PnBwdMask(Function):
    """dz = lrelu'(a) * pixel_norm_backward(a, r, dy) in one pass.  With M = lrelu'(a) (piecewise constant)
    and J(a) the symmetric pixel-norm Jacobian: dz = M J dy, so for an incoming u
    d/d(dy) = J (M u) and d/da = pn_bwd2(a, r, dy, M u) -- returned as M * (d/da), the gradient w.r.t. the
    pre-activation, because `a` only ever comes from a premasked ConvLayer."""

    @staticmethod
    def forward(ctx, a, r, dy):
        ctx.save_for_backward(a, r, dy)
        return K.pn_bwd_mask(a, r, dy, False)[0]

    @staticmethod
    @once_differentiable
    def backward(ctx, u):
        a, r, dy = ctx.saved_tensors
        # `a` is the output of a premasked ConvLayer: every gradient handed to it must already carry the mask
        ga, gdy = K.pn_bwd_mask_second(a, r, dy, u)
        return ga, None, gdy


class LeakyRelu(Function):
    @staticmethod
    def forward(ctx, x):
        y = K.lrelu(x)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        return MaskMul.apply(g, y)


class ColSum(Function):
    """[..., C] -> [C] (bias gradient)."""

    @staticmethod
    def forward(ctx, v):
        ctx.lead = tuple(v.shape[:-1])
        return K.col_sum(v)

    @staticmethod
    def backward(ctx, g):
        return RowBroadcast.apply(g, ctx.lead)


class RowBroadcast(Function):
    @staticmethod
    def forward(ctx, s, lead):
        return K.row_broadcast(s, lead)

    @staticmethod
    def backward(ctx, g):
        return ColSum.apply(g), None


class BiasAct(Function):
    """act(x + bias) on [..., C] (dense layers)."""

    @staticmethod
    def forward(ctx, x, bias, act):
        y = K.bias_act(x, bias, act)
        ctx.act = act
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dz = MaskMul.apply(dy, y) if ctx.act == ACT_LRELU else dy
        db = ColSum.apply(dz) if (ctx.needs_input_grad[1] and not _SKIP_WGRAD) else None
        return dz, db, None


class Tanh(Function):
    @staticmethod
    def forward(ctx, x):
        y = K.tanh_fwd(x)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return TanhBwd.apply(y, dy)


class TanhBwd(Function):
    """dy * (1 - y^2); d/dy-out = -2 y dy u, d/d(dy) = u (1 - y^2)."""

    @staticmethod
    def forward(ctx, y, dy):
        ctx.save_for_backward(y, dy)
        return K.tanh_bwd(y, dy)

    @staticmethod
    @once_differentiable
    def backward(ctx, u):
        y, dy = ctx.saved_tensors
        return K.tanh_bwd2(y, dy, u), K.tanh_bwd(y, u)


class Axpby(Function):
    """alpha * a + beta * b (lerp, networks.py:10-11)."""

    @staticmethod
    def forward(ctx, a, b, alpha, beta):
        ctx.ab = (alpha, beta)
        return K.axpby(a, b, alpha, beta)

    @staticmethod
    def backward(ctx, g):
        alpha, beta = ctx.ab
        return Scale.apply(g, alpha), Scale.apply(g, beta), None, None


class AxpbyDev(Function):
    """coef[0] * a + coef[1] * b with `coef` a DEVICE tensor [2]: the progressive-growing blend (networks.py:10-11,
    126-152) inside a replayed CUDA graph, where the weight depends on global_step and cannot be a launch argument."""

    @staticmethod
    def forward(ctx, a, b, coef):
        ctx.coef = coef
        return K.axpby_dev(a, b, coef, 0, 1)

    @staticmethod
    def backward(ctx, g):
        return ScaleDev.apply(g, ctx.coef, 0), ScaleDev.apply(g, ctx.coef, 1), None


class ScaleDev(Function):
    @staticmethod
    def forward(ctx, a, coef, idx):
        ctx.coef, ctx.idx = coef, idx
        return K.axpby_dev(a, None, coef, idx, idx)

    @staticmethod
    def backward(ctx, g):
        return ScaleDev.apply(g, ctx.coef, ctx.idx), None, None


class Scale(Function):
    @staticmethod
    def forward(ctx, a, s):
        ctx.s = s
        return K.axpby(a, None, s, 0.0)

    @staticmethod
    def backward(ctx, g):
        return Scale.apply(g, ctx.s), None


# ----------------------------------------------------------------------------- pixel norm
class PixelNorm(Function):
    """ops.py:330-333 over the last (channel) axis."""

    @staticmethod
    def forward(ctx, a, eps):
        y, r = K.pn_fwd(a, eps)
        ctx.save_for_backward(a, r)
        return y

    @staticmethod
    def backward(ctx, dy):
        a, r = ctx.saved_tensors
        return PixelNormBwd.apply(a, r, dy), None


class PixelNormOfLayer(Function):
    """Pixel normalisation of the output `a` of a leaky-relu ConvLayer built with premasked=True (the
    generator's conv -> leaky_relu -> pixel_normalization triple, networks.py:57-68, 82-93).  Its backward
    returns the gradient w.r.t. the layer's PRE-activation: pixel-norm backward and the leaky-relu mask
    (taken from the sign of a) share one pass."""

    @staticmethod
    def forward(ctx, a, eps):
        y, r = K.pn_fwd(a, eps)
        ctx.save_for_backward(a, r)
        return y

    @staticmethod
    def backward(ctx, dy):
        a, r = ctx.saved_tensors
        return PnBwdMask.apply(a, r, dy), None


class PixelNormBwd(Function):
    """da = r dy - (r^3/C)(a.dy) a.  Symmetric in the sense J(a)^T = J(a), so d/d(dy) reuses itself;
    d/da is the hand-derived second-order kernel (gs_pixel_norm_bwd2)."""

    @staticmethod
    def forward(ctx, a, r, dy):
        ctx.save_for_backward(a, r, dy)
        return K.pn_bwd(a, r, dy)

    @staticmethod
    @once_differentiable
    def backward(ctx, u):
        a, r, dy = ctx.saved_tensors
        return K.pn_bwd2(a, r, dy, u), None, K.pn_bwd(a, r, u)


# ----------------------------------------------------------------------------- dense trio / embedding
class DenseF(Function):
    @staticmethod
    def forward(ctx, x, w, alpha):
        ctx.alpha = alpha
        ctx.save_for_backward(x, w)
        return K.dense_fwd(x, w, alpha)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = DenseD.apply(dy, w, ctx.alpha) if ctx.needs_input_grad[0] else None
        dw = DenseW.apply(x, dy, ctx.alpha) if (ctx.needs_input_grad[1] and not _SKIP_WGRAD) else None
        return dx, dw, None


class DenseD(Function):
    @staticmethod
    def forward(ctx, dy, w, alpha):
        ctx.alpha = alpha
        ctx.save_for_backward(dy, w)
        return K.dense_dgrad(dy, w, alpha)

    @staticmethod
    def backward(ctx, g):
        dy, w = ctx.saved_tensors
        gdy = DenseF.apply(g, w, ctx.alpha) if ctx.needs_input_grad[0] else None
        gw = DenseW.apply(g, dy, ctx.alpha) if ctx.needs_input_grad[1] else None
        return gdy, gw, None


class DenseW(Function):
    @staticmethod
    def forward(ctx, x, dy, alpha):
        ctx.alpha = alpha
        ctx.save_for_backward(x, dy)
        return K.dense_wgrad(x, dy, alpha)

    @staticmethod
    def backward(ctx, gw):
        x, dy = ctx.saved_tensors
        gx = DenseD.apply(dy, gw, ctx.alpha) if ctx.needs_input_grad[0] else None
        gdy = DenseF.apply(x, gw, ctx.alpha) if ctx.needs_input_grad[1] else None
        return gx, gdy, None


class Embedding(Function):
    @staticmethod
    def forward(ctx, table, idx, alpha):
        ctx.alpha, ctx.rows = alpha, table.shape[0]
        ctx.save_for_backward(idx)
        return K.embedding_fwd(table, idx, alpha)

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        if _SKIP_WGRAD or not ctx.needs_input_grad[0]:
            return None, None, None
        return EmbeddingBwd.apply(dy, idx, ctx.rows, ctx.alpha), None, None


class EmbeddingBwd(Function):
    @staticmethod
    def forward(ctx, dy, idx, rows, alpha):
        ctx.alpha = alpha
        ctx.save_for_backward(idx)
        return K.embedding_bwd(dy, idx, rows, alpha)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        return Embedding.apply(g, idx, ctx.alpha), None, None, None


# ----------------------------------------------------------------------------- resampling / layout
class Upscale(Function):
    """scale * nearest-neighbour repeat (ops.py:283-291)."""

    @staticmethod
    def forward(ctx, x, fh, fw, scale):
        ctx.cfg = (fh, fw, scale)
        return K.upscale(x, fh, fw, scale)

    @staticmethod
    def backward(ctx, g):
        return Pool.apply(g, *ctx.cfg), None, None, None


class Pool(Function):
    """scale * sum-pool with kernel = stride = (fh, fw); downscale2d (ops.py:294-305) is scale 1/(fh*fw)."""

    @staticmethod
    def forward(ctx, x, fh, fw, scale):
        ctx.cfg = (fh, fw, scale)
        return K.pool(x, fh, fw, scale)

    @staticmethod
    def backward(ctx, g):
        return Upscale.apply(g, *ctx.cfg), None, None, None


class TransposeInner(Function):
    """[n, a, b] -> [n, b, a]; NCHW <-> NHWC with (a, b) = (C, H*W)."""

    @staticmethod
    def forward(ctx, x):
        return K.transpose_inner(x)

    @staticmethod
    def backward(ctx, g):
        return TransposeInner.apply(g)


def nchw_to_nhwc(x):
    n, c, h, w = x.shape
    return TransposeInner.apply(x.reshape(n, c, h * w)).reshape(n, h, w, c)


def nhwc_to_nchw(x):
    n, h, w, c = x.shape
    return TransposeInner.apply(x.reshape(n, h * w, c)).reshape(n, c, h, w)


# ----------------------------------------------------------------------------- per-sample reductions
class RowDot(Function):
    """[B, E] x [B, E] -> [B]."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return K.row_dot(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        ga = RowScale.apply(b, g) if ctx.needs_input_grad[0] else None
        gb = RowScale.apply(a, g) if ctx.needs_input_grad[1] else None
        return ga, gb


class RowScale(Function):
    """out[b, e] = a[b, e] * s[b]."""

    @staticmethod
    def forward(ctx, a, s):
        ctx.save_for_backward(a, s)
        return K.row_scale(a, s, 1.0)

    @staticmethod
    def backward(ctx, u):
        a, s = ctx.saved_tensors
        ga = RowScale.apply(u, s) if ctx.needs_input_grad[0] else None
        gs = RowDot.apply(u, a) if ctx.needs_input_grad[1] else None
        return ga, gs


# ----------------------------------------------------------------------------- minibatch stddev
class BatchStddev(Function):
    """x [B, E] -> statistic [B/groups] (ops.py:336-347 up to the final tile)."""

    @staticmethod
    def forward(ctx, x, groups, eps):
        ctx.cfg = (groups, eps)
        ctx.save_for_backward(x)
        return K.stddev_fwd(x, groups, eps)

    @staticmethod
    def backward(ctx, df):
        (x,) = ctx.saved_tensors
        return BatchStddevBwd.apply(x, df, *ctx.cfg), None, None


class BatchStddevBwd(Function):
    @staticmethod
    def forward(ctx, x, df, groups, eps):
        ctx.cfg = (groups, eps)
        ctx.save_for_backward(x, df)
        return K.stddev_bwd(x, df, groups, eps)

    @staticmethod
    @once_differentiable
    def backward(ctx, u):
        x, df = ctx.saved_tensors
        gx, q = K.stddev_bwd2(x, df, u, *ctx.cfg)
        return gx, q, None, None


# ----------------------------------------------------------------------------- pitch classifier (first-order gradients)
class GroupNorm(Function):
    """group_normalization (ops.py:118-146) [+ relu] on NHWC; the backward is a kernel (first order only: nothing in the
    classifier's training differentiates twice)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps, relu):
        y, stats = K.group_norm(x, gamma, beta, groups, eps, relu)
        ctx.cfg = (groups, eps, relu)
        ctx.save_for_backward(x, y, stats, gamma)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, y, stats, gamma = ctx.saved_tensors
        groups, eps, relu = ctx.cfg
        dx, dgamma, dbeta = K.group_norm_bwd(x, y, dy.contiguous(), stats, gamma, groups, eps, relu)
        return dx, dgamma, dbeta, None, None, None


class MaxPool(Function):
    @staticmethod
    def forward(ctx, x, ksize, stride):
        y = K.max_pool(x, ksize, stride)
        ctx.cfg = (ksize, stride)
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        return K.max_pool_bwd(x, y, dy.contiguous(), *ctx.cfg), None, None


class SpatialMean(Function):
    @staticmethod
    def forward(ctx, x):
        ctx.shape = tuple(x.shape)
        return K.spatial_mean(x)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        return K.spatial_mean_bwd(dy.contiguous(), ctx.shape)
