"""Thin tensor-level wrappers over the C ABI: allocate outputs with torch (device memory plumbing),
pass raw device pointers and the current CUDA stream to libgansynth_b200.so.

Activations are NHWC fp32.  Every method requires CUDA tensors and fails loudly otherwise -- there is
no CPU path in the product.
"""
import os

import torch

from . import _lib

IMPL_AUTO, IMPL_NAIVE, IMPL_TILED, IMPL_TC, IMPL_FP32 = 0, 1, 2, 3, 4
EPI_NONE, EPI_MASK, EPI_PIXEL_NORM = 0, 1, 2


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(*tensors):
    out = []
    for t in tensors:
        if t is None:
            out.append(None)
            continue
        if not t.is_cuda:
            raise _lib.GansynthLibraryError("gansynth_b200 kernels need CUDA tensors (got %s); no CPU fallback" % t.device)
        if t.dtype != torch.float32:
            raise TypeError("expected float32, got %s" % t.dtype)
        if not t.is_contiguous():
            t = t.contiguous()
        if t.data_ptr() % 16:
            t = t.clone()   # the float4 paths of the kernels need 16-byte aligned base pointers
        out.append(t)
    return out


class CudaBackend(object):
    """The primitive-kernel API the autograd functions in functional.py are written against."""

    impl = IMPL_AUTO
    # implementation used for LAYER FORWARD convolutions (`precise=True`): their outputs decide the
    # leaky-relu masks, so GS_CONV_FWD_IMPL=4 routes them to the exact-fp32 kernels while the gradient
    # forms stay on the tensor cores (default -1: same as every other convolution)
    fwd_impl = int(os.environ.get("GS_CONV_FWD_IMPL", "-1"))

    # address ranges of the packed parameter buffers: convolution weights inside them are flagged as cacheable
    param_ranges = ()

    def register_parameters(self, flats):
        self.param_ranges = tuple((t.data_ptr(), t.data_ptr() + t.numel() * t.element_size()) for t in flats)
        self.weight_cache_reset()

    def weight_cache_reset(self):
        """Drops every cached split weight (new parameter buffers: the cached addresses mean nothing any more)."""
        _lib.reset_weight_cache()

    def weight_cache_refresh(self, flat=None):
        """Re-splits in place the cached copies of the parameters inside `flat` (None: all) -- to be called after every
        change of parameter VALUES (optimiser update, load): the cache then never holds a stale copy, its slots keep
        their addresses (CUDA graphs stay valid), and the sub-steps contain no split kernels."""
        if not _lib.is_loaded() or not _lib._contexts:
            return
        if flat is None:
            _lib.call("gs_conv_weight_cache_refresh", None, None, _stream())
        else:
            _lib.call("gs_conv_weight_cache_refresh", flat.data_ptr(), flat.data_ptr() + flat.numel() * flat.element_size(),
                      _stream())

    def _impl(self, precise, w=None):
        impl = self.fwd_impl if (precise and self.fwd_impl >= 0) else self.impl
        if w is not None and self.param_ranges:
            p = w.data_ptr()
            if any(lo <= p < hi for lo, hi in self.param_ranges):
                impl |= 0x100      # GS_IMPL_PARAM_WEIGHT
        return impl

    # ------------------------------------------------------------------ convolution family
    def conv_c(self, x, w, bias, ksize, stride, wswap, alpha, act, precise=False, mask_src=None):
        """Gather convolution; with `mask_src` (shaped like the output) the result is multiplied by lrelu'(mask_src)
        in the kernel epilogue (GS_EPI_MASK)."""
        return self._conv("gs_conv2d_fwd_ex", x, w, bias, ksize, stride, wswap, alpha, act, precise,
                          EPI_MASK if mask_src is not None else EPI_NONE, mask_src, None, 0.0)

    def conv_t(self, dy, w, bias, ksize, stride, wswap, alpha, act, precise=False, mask_src=None):
        """Transposed (input-gradient form) convolution; `mask_src` as in conv_c."""
        return self._conv("gs_conv2d_dgrad_ex", dy, w, bias, ksize, stride, wswap, alpha, act, precise,
                          EPI_MASK if mask_src is not None else EPI_NONE, mask_src, None, 0.0)

    def conv_pn(self, x, w, bias, form, ksize, stride, wswap, alpha, eps):
        """pixel_normalization(leaky_relu(alpha * conv(x, w) + bias)) in one kernel (GS_EPI_PIXEL_NORM), conv in gather
        ('c') or transposed ('t') form -> (y, r) with r = 1 / sqrt(mean_c(a^2) + eps) per pixel."""
        name = "gs_conv2d_fwd_ex" if form == "c" else "gs_conv2d_dgrad_ex"
        return self._conv(name, x, w, bias, ksize, stride, wswap, alpha, 1, True, EPI_PIXEL_NORM, None, True, eps)

    def _conv(self, name, x, w, bias, ksize, stride, wswap, alpha, act, precise, epi, aux, want_r, eps):
        x, w, bias, aux = _chk(x, w, bias, aux)
        n = x.shape[0]
        if name == "gs_conv2d_fwd_ex":
            h, wd, ci = x.shape[1:]
            co = w.shape[2] if wswap else w.shape[3]
            assert (w.shape[3] if wswap else w.shape[2]) == ci, "conv_c: weight/input channel mismatch"
            out = torch.empty((n, h // stride, wd // stride, co), device=x.device, dtype=torch.float32)
        else:
            oh, ow, co = x.shape[1:]
            ci = w.shape[3] if wswap else w.shape[2]
            assert (w.shape[2] if wswap else w.shape[3]) == co, "conv_t: weight/input channel mismatch"
            h, wd = oh * stride, ow * stride
            out = torch.empty((n, h, wd, ci), device=x.device, dtype=torch.float32)
        if aux is not None:
            assert aux.shape == out.shape, "conv: mask source must be shaped like the output"
        r = torch.empty(out.shape[:-1], device=x.device, dtype=torch.float32) if want_r else None
        _lib.call(name, _ptr(x), _ptr(w), _ptr(bias), _ptr(out), n, h, wd, ci, co, ksize, stride, int(wswap),
                  float(alpha), int(act), int(epi), _ptr(aux), _ptr(r), float(eps), self._impl(precise, w), _stream())
        return (out, r) if want_r else out

    def conv_w(self, x, dy, ksize, stride, wswap, alpha, bias_of=None, out=None):
        """Filter gradient; with `bias_of` ('dy' or 'x') also the column sum of that operand (the bias gradient of the
        layer, taken from the same pass) -> (dw, db).  `out` = (dw buffer, db buffer or None): the results are ADDED to
        those tensors (gradient accumulation in the caller's buffer) and nothing is returned."""
        x, dy = _chk(x, dy)
        n, h, wd, ci = x.shape
        co = dy.shape[3]
        shape = (ksize, ksize, co, ci) if wswap else (ksize, ksize, ci, co)
        if out is not None:
            dw, db = out
            assert tuple(dw.shape) == shape and dw.is_contiguous() and (db is None or bias_of in ("dy", "x"))
            _lib.call("gs_conv2d_wgrad_ex", _ptr(x), _ptr(dy), _ptr(dw), _ptr(db), int(bias_of == "x"), 1, n, h, wd, ci, co,
                      ksize, stride, int(wswap), float(alpha), self.impl, _stream())
            return None
        dw = torch.empty(shape, device=x.device, dtype=torch.float32)
        if bias_of is None:
            _lib.call("gs_conv2d_wgrad", _ptr(x), _ptr(dy), _ptr(dw), n, h, wd, ci, co, ksize, stride, int(wswap),
                      float(alpha), self.impl, _stream())
            return dw
        assert bias_of in ("dy", "x")
        db = torch.empty((ci if bias_of == "x" else co,), device=x.device, dtype=torch.float32)
        _lib.call("gs_conv2d_wgrad_ex", _ptr(x), _ptr(dy), _ptr(dw), _ptr(db), int(bias_of == "x"), 0, n, h, wd, ci, co, ksize,
                  stride, int(wswap), float(alpha), self.impl, _stream())
        return dw, db

    # ------------------------------------------------------------------ dense / embedding
    def dense_fwd(self, x, w, alpha):
        x, w = _chk(x, w)
        m, k = x.shape
        n = w.shape[1]
        y = torch.empty((m, n), device=x.device, dtype=torch.float32)
        _lib.call("gs_dense_fwd", _ptr(x), _ptr(w), _ptr(y), m, k, n, float(alpha), _stream())
        return y

    def dense_dgrad(self, dy, w, alpha):
        dy, w = _chk(dy, w)
        m, n = dy.shape
        k = w.shape[0]
        dx = torch.empty((m, k), device=dy.device, dtype=torch.float32)
        _lib.call("gs_dense_dgrad", _ptr(dy), _ptr(w), _ptr(dx), m, k, n, float(alpha), _stream())
        return dx

    def dense_wgrad(self, x, dy, alpha):
        x, dy = _chk(x, dy)
        m, k = x.shape
        n = dy.shape[1]
        dw = torch.empty((k, n), device=x.device, dtype=torch.float32)
        _lib.call("gs_dense_wgrad", _ptr(x), _ptr(dy), _ptr(dw), m, k, n, float(alpha), _stream())
        return dw

    def embedding_fwd(self, table, idx, alpha):
        (table,) = _chk(table)
        idx = idx.contiguous()
        assert idx.dtype == torch.int64 and idx.is_cuda
        b, units = idx.shape[0], table.shape[1]
        out = torch.empty((b, units), device=table.device, dtype=torch.float32)
        _lib.call("gs_embedding_fwd", _ptr(table), _ptr(idx), _ptr(out), b, units, float(alpha), _stream())
        return out

    def embedding_bwd(self, dy, idx, rows, alpha):
        (dy,) = _chk(dy)
        idx = idx.contiguous()
        b, units = dy.shape
        out = torch.empty((rows, units), device=dy.device, dtype=torch.float32)
        _lib.call("gs_embedding_bwd", _ptr(dy), _ptr(idx), _ptr(out), b, rows, units, float(alpha), _stream())
        return out

    # ------------------------------------------------------------------ elementwise
    def _ew(self, name, *ins, extra=()):
        ins = _chk(*ins)
        out = torch.empty_like(ins[0])
        _lib.call(name, *[_ptr(t) for t in ins], _ptr(out), *extra, ins[0].numel(), _stream())
        return out

    def lrelu(self, x):
        return self._ew("gs_lrelu", x)

    def mask_mul(self, v, y):
        return self._ew("gs_lrelu_mask_mul", v, y)

    def mask_mul_colsum(self, v, y):
        """(v * lrelu'(y), column sums of that product over the last axis) in one pass."""
        v, y = _chk(v, y)
        c = v.shape[-1]
        if c % 4 or c > 256 or 256 % (c // 4):
            out = self.mask_mul(v, y)
            return out, self.col_sum(out)
        out = torch.empty_like(v)
        cs = torch.empty((c,), device=v.device, dtype=torch.float32)
        _lib.call("gs_lrelu_mask_mul_colsum", _ptr(v), _ptr(y), _ptr(out), _ptr(cs), v.numel() // c, c, _stream())
        return out, cs

    def tanh_fwd(self, x):
        return self._ew("gs_tanh_fwd", x)

    def tanh_bwd(self, y, dy):
        return self._ew("gs_tanh_bwd", y, dy)

    def tanh_bwd2(self, y, dy, u):
        return self._ew("gs_tanh_bwd2", y, dy, u)

    def axpby(self, a, b, alpha, beta):
        a, b = _chk(a, b)
        out = torch.empty_like(a)
        _lib.call("gs_axpby", _ptr(a), _ptr(b), _ptr(out), float(alpha), float(beta), a.numel(), _stream())
        return out

    def axpby_dev(self, a, b, coef, ia, ib):
        a, b, coef = _chk(a, b, coef)
        out = torch.empty_like(a)
        _lib.call("gs_axpby_dev", _ptr(a), _ptr(b), _ptr(out), _ptr(coef), int(ia), int(ib), a.numel(), _stream())
        return out

    def bias_act(self, x, bias, act):
        x, bias = _chk(x, bias)
        c = x.shape[-1]
        out = torch.empty_like(x)
        _lib.call("gs_bias_act", _ptr(x), _ptr(bias), _ptr(out), x.numel() // c, c, int(act), _stream())
        return out

    def row_broadcast(self, s, lead_shape):
        (s,) = _chk(s)
        c = s.numel()
        out = torch.empty(tuple(lead_shape) + (c,), device=s.device, dtype=torch.float32)
        _lib.call("gs_row_broadcast", _ptr(s), _ptr(out), out.numel() // c, c, _stream())
        return out

    def col_sum(self, v):
        (v,) = _chk(v)
        c = v.shape[-1]
        out = torch.empty((c,), device=v.device, dtype=torch.float32)
        _lib.call("gs_col_sum", _ptr(v), _ptr(out), v.numel() // c, c, _stream())
        return out

    # ------------------------------------------------------------------ pixel norm
    def pn_fwd(self, a, eps):
        (a,) = _chk(a)
        c = a.shape[-1]
        y = torch.empty_like(a)
        r = torch.empty(a.shape[:-1], device=a.device, dtype=torch.float32)
        _lib.call("gs_pixel_norm_fwd", _ptr(a), _ptr(y), _ptr(r), a.numel() // c, c, float(eps), _stream())
        return y, r

    def pn_bwd(self, a, r, dy):
        a, r, dy = _chk(a, r, dy)
        c = a.shape[-1]
        da = torch.empty_like(a)
        _lib.call("gs_pixel_norm_bwd", _ptr(a), _ptr(r), _ptr(dy), _ptr(da), a.numel() // c, c, _stream())
        return da

    def pn_bwd_mask(self, a, r, dy, want_colsum):
        """lrelu'(a) * pixel-norm backward (a, r, dy); with want_colsum also the column sums (bias gradient)."""
        a, r, dy = _chk(a, r, dy)
        c = a.shape[-1]
        if c % 4 or c > 256:
            dz = self.mask_mul(self.pn_bwd(a, r, dy), a)
            return dz, (self.col_sum(dz) if want_colsum else None)
        dz = torch.empty_like(a)
        cs = torch.empty((c,), device=a.device, dtype=torch.float32) if want_colsum else None
        _lib.call("gs_pixel_norm_bwd_mask", _ptr(a), _ptr(r), _ptr(dy), _ptr(dz), _ptr(cs), a.numel() // c, c, _stream())
        return dz, cs

    @staticmethod
    def _pn_vec_ok(c):
        v = c // 4
        return c % 4 == 0 and ((v <= 32 and v & (v - 1) == 0) or v in (64, 128))

    def pn_bwd_mask_second(self, a, r, dy, u):
        """Second-order pieces of pn_bwd_mask for an incoming u (M = lrelu'(a)):
        (M * pn_bwd2(a, r, dy, M u), pn_bwd(a, r, M u)) without materialising M u."""
        a, r, dy, u = _chk(a, r, dy, u)
        c = a.shape[-1]
        if not self._pn_vec_ok(c):
            mu = self.mask_mul(u, a)
            return self.mask_mul(self.pn_bwd2(a, r, dy, mu), a), self.pn_bwd(a, r, mu)
        ga, gdy = torch.empty_like(a), torch.empty_like(a)
        rows = a.numel() // c
        _lib.call("gs_pixel_norm_bwd2_masked", _ptr(a), _ptr(r), _ptr(dy), _ptr(u), _ptr(ga), rows, c, _stream())
        _lib.call("gs_pixel_norm_bwd_premask", _ptr(a), _ptr(r), _ptr(u), _ptr(gdy), rows, c, _stream())
        return ga, gdy

    def pn_bwd_mask_y(self, y, r, dy, want_colsum=False):
        """pn_bwd_mask for a fused conv + pixel-norm layer that kept (y = a * r, r): lrelu'(y) * pixel-norm backward."""
        y, r, dy = _chk(y, r, dy)
        c = y.shape[-1]
        dz = torch.empty_like(y)
        cs = torch.empty((c,), device=y.device, dtype=torch.float32) if want_colsum else None
        _lib.call("gs_pixel_norm_bwd_mask_y", _ptr(y), _ptr(r), _ptr(dy), _ptr(dz), _ptr(cs), y.numel() // c, c, _stream())
        return dz, cs

    def pn_bwd_mask_second_y(self, y, r, dy, u):
        """pn_bwd_mask_second in the (y, r) form."""
        y, r, dy, u = _chk(y, r, dy, u)
        c = y.shape[-1]
        ga, gdy = torch.empty_like(y), torch.empty_like(y)
        rows = y.numel() // c
        _lib.call("gs_pixel_norm_bwd2_pair_y", _ptr(y), _ptr(r), _ptr(dy), _ptr(u), _ptr(ga), _ptr(gdy), rows, c, _stream())
        return ga, gdy

    def pn_bwd2(self, a, r, dy, u):
        a, r, dy, u = _chk(a, r, dy, u)
        c = a.shape[-1]
        ga = torch.empty_like(a)
        _lib.call("gs_pixel_norm_bwd2", _ptr(a), _ptr(r), _ptr(dy), _ptr(u), _ptr(ga), a.numel() // c, c, _stream())
        return ga

    # ------------------------------------------------------------------ minibatch stddev on [B, E]
    def stddev_fwd(self, x, groups, eps):
        (x,) = _chk(x)
        b, e = x.shape
        stat = torch.empty((b // groups,), device=x.device, dtype=torch.float32)
        _lib.call("gs_batch_stddev_fwd", _ptr(x), _ptr(stat), b, e, groups, float(eps), _stream())
        return stat

    def stddev_bwd(self, x, df, groups, eps):
        x, df = _chk(x, df)
        b, e = x.shape
        dx = torch.empty_like(x)
        _lib.call("gs_batch_stddev_bwd", _ptr(x), _ptr(df), _ptr(dx), b, e, groups, float(eps), _stream())
        return dx

    def stddev_bwd2(self, x, df, u, groups, eps):
        x, df, u = _chk(x, df, u)
        b, e = x.shape
        gx = torch.empty_like(x)
        q = torch.empty((b // groups,), device=x.device, dtype=torch.float32)
        _lib.call("gs_batch_stddev_bwd2", _ptr(x), _ptr(df), _ptr(u), _ptr(gx), _ptr(q), b, e, groups, float(eps),
                  _stream())
        return gx, q

    # ------------------------------------------------------------------ resampling / layout
    def upscale(self, x, fh, fw, scale):
        (x,) = _chk(x)
        n, h, w, c = x.shape
        out = torch.empty((n, h * fh, w * fw, c), device=x.device, dtype=torch.float32)
        _lib.call("gs_upscale2d", _ptr(x), _ptr(out), n, h, w, c, fh, fw, float(scale), _stream())
        return out

    def pool(self, x, fh, fw, scale):
        (x,) = _chk(x)
        n, hh, ww, c = x.shape
        h, w = hh // fh, ww // fw
        out = torch.empty((n, h, w, c), device=x.device, dtype=torch.float32)
        _lib.call("gs_pool2d", _ptr(x), _ptr(out), n, h, w, c, fh, fw, float(scale), _stream())
        return out

    # ------------------------------------------------------------------ pitch classifier
    def group_norm(self, x, gamma, beta, groups, eps, relu):
        """ops.py:118-146 on NHWC, optionally followed by relu (networks.py:318-322) -> (y, stats [n, groups, 2])."""
        x, gamma, beta = _chk(x, gamma, beta)
        n, c = x.shape[0], x.shape[-1]
        hw = x.numel() // (n * c)
        y = torch.empty_like(x)
        stats = torch.empty((n, groups, 2), device=x.device, dtype=torch.float32)
        _lib.call("gs_group_norm_fwd", _ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(stats), n, hw, c, int(groups), float(eps),
                  int(bool(relu)), _stream())
        return y, stats

    def group_norm_bwd(self, x, y, dy, stats, gamma, groups, eps, relu):
        x, y, dy, stats, gamma = _chk(x, y, dy, stats, gamma)
        n, c = x.shape[0], x.shape[-1]
        hw = x.numel() // (n * c)
        dx = torch.empty_like(x)
        dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(gamma)
        red = torch.empty((n, groups, 2), device=x.device, dtype=torch.float32)
        _lib.call("gs_group_norm_bwd", _ptr(x), _ptr(y), _ptr(dy), _ptr(stats), _ptr(gamma), _ptr(dx), _ptr(dgamma), _ptr(dbeta),
                  _ptr(red), n, hw, c, int(groups), float(eps), int(bool(relu)), _stream())
        return dx, dgamma, dbeta

    def max_pool(self, x, ksize, stride):
        (x,) = _chk(x)
        n, h, w, c = x.shape
        y = torch.empty((n, -(-h // stride), -(-w // stride), c), device=x.device, dtype=torch.float32)
        _lib.call("gs_max_pool2d", _ptr(x), _ptr(y), n, h, w, c, int(ksize), int(stride), _stream())
        return y

    def max_pool_bwd(self, x, y, dy, ksize, stride):
        x, y, dy = _chk(x, y, dy)
        n, h, w, c = x.shape
        dx = torch.empty_like(x)
        _lib.call("gs_max_pool2d_bwd", _ptr(x), _ptr(y), _ptr(dy), _ptr(dx), n, h, w, c, int(ksize), int(stride), _stream())
        return dx

    def spatial_mean(self, x):
        (x,) = _chk(x)
        n, c = x.shape[0], x.shape[-1]
        y = torch.empty((n, c), device=x.device, dtype=torch.float32)
        _lib.call("gs_spatial_mean", _ptr(x), _ptr(y), n, x.numel() // (n * c), c, _stream())
        return y

    def spatial_mean_bwd(self, dy, shape):
        (dy,) = _chk(dy)
        n, c = dy.shape
        dx = torch.empty(tuple(shape), device=dy.device, dtype=torch.float32)
        _lib.call("gs_spatial_mean_bwd", _ptr(dy), _ptr(dx), n, dx.numel() // (n * c), c, _stream())
        return dx

    def momentum_step(self, p, g, accum, wd, lr, momentum, nesterov, grad_scale=1.0):
        """tf.train.MomentumOptimizer on flat buffers (gs_momentum_step); wd: per-element L2 coefficient or None."""
        for tns in (p, g, accum):
            assert tns.is_cuda and tns.is_contiguous() and tns.dtype == torch.float32
        _lib.call("gs_momentum_step", _ptr(p), _ptr(g), _ptr(accum), _ptr(wd), p.numel(), float(lr), float(momentum),
                  int(bool(nesterov)), float(grad_scale), _stream())

    def transpose_inner(self, x):
        """[n, a, b] -> [n, b, a]"""
        (x,) = _chk(x)
        n, a, b = x.shape
        out = torch.empty((n, b, a), device=x.device, dtype=torch.float32)
        _lib.call("gs_transpose_inner", _ptr(x), _ptr(out), n, a, b, _stream())
        return out

    # ------------------------------------------------------------------ per-sample row ops on [B, E]
    def row_dot(self, a, b):
        a, b = _chk(a, b)
        rows, e = a.shape
        out = torch.empty((rows,), device=a.device, dtype=torch.float32)
        _lib.call("gs_row_dot", _ptr(a), _ptr(b), _ptr(out), rows, e, _stream())
        return out

    def row_scale(self, a, s, alpha=1.0):
        a, s = _chk(a, s)
        rows, e = a.shape
        out = torch.empty_like(a)
        _lib.call("gs_row_scale", _ptr(a), _ptr(s), _ptr(out), rows, e, float(alpha), _stream())
        return out

    # ------------------------------------------------------------------ optimiser
    def adam_step(self, p, g, m, v, lr, beta1, beta2, eps, t, grad_scale=1.0):
        for tns in (p, g, m, v):
            assert tns.is_cuda and tns.is_contiguous() and tns.dtype == torch.float32
        _lib.call("gs_adam_step", _ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), float(lr), float(beta1),
                  float(beta2), float(eps), int(t), float(grad_scale), _stream())

    def adam_slice(self, n, rank, world):
        import ctypes
        lo, hi = ctypes.c_longlong(), ctypes.c_longlong()
        _lib.host_call("gs_adam_slice", int(n), int(rank), int(world), ctypes.byref(lo), ctypes.byref(hi))
        return int(lo.value), int(hi.value)

    def adam_step_allreduce(self, p_local, m, v, grad_mc, param_mc, grad_peers, param_peers, rank, world, lr, beta1, beta2,
                            eps, t, grad_scale):
        """Gradient all-reduce fused with TF-Adam over NVLink (gs_adam_step_allreduce): `grad_mc` / `param_mc` are the
        NVSwitch multicast addresses (ints, 0 = none), `grad_peers` / `param_peers` int64 CUDA tensors holding the peers'
        pointers.  The caller brackets the call with cross-rank barriers."""
        for tns in (p_local, m, v):
            assert tns.is_cuda and tns.is_contiguous() and tns.dtype == torch.float32
        _lib.call("gs_adam_step_allreduce", _ptr(p_local), _ptr(m), _ptr(v), int(grad_mc) or None, int(param_mc) or None,
                  _ptr(grad_peers), _ptr(param_peers), p_local.numel(), int(rank), int(world), float(lr), float(beta1),
                  float(beta2), float(eps), int(t), float(grad_scale), _stream())

    # ------------------------------------------------------------------ spectral
    def spectrogram_fwd(self, wave, consts, time_steps, frames_per_run):
        (wave,) = _chk(wave)
        b, wave_len = wave.shape
        logmel = torch.empty((b, time_steps, 1024), device=wave.device, dtype=torch.float32)
        inst = torch.empty_like(logmel)
        runs = -(-time_steps // frames_per_run)
        scratch = torch.empty((b, runs, 1024), device=wave.device, dtype=torch.float32) if runs > 1 else None
        _lib.call("gs_spectrogram_fwd", _ptr(wave), _ptr(consts["hann"]), _ptr(consts["mel_k0"]), _ptr(consts["mel_w"]),
                  _ptr(logmel), _ptr(inst), _ptr(scratch), b, wave_len, time_steps,
                  frames_per_run, _stream())
        return logmel, inst

    def spectrogram_generic(self, wave, consts, time_steps, bins, frame_step):
        """convert_to_spectrogram for any (bins, overlap): csrc/spectral_generic.cu."""
        (wave,) = _chk(wave)
        b, wave_len = wave.shape
        logmel = torch.empty((b, time_steps, bins), device=wave.device, dtype=torch.float32)
        inst = torch.empty_like(logmel)
        scratch = torch.empty_like(logmel)
        _lib.call("gs_spectrogram_generic", _ptr(wave), _ptr(consts["hann"]), _ptr(consts["mel"]), _ptr(logmel), _ptr(inst),
                  _ptr(scratch), b, wave_len, time_steps, int(bins), int(frame_step), _stream())
        return logmel, inst

    def waveform_generic(self, logmel, inst, consts, wave_len, bins, frame_step):
        """convert_to_waveform for any (bins, overlap)."""
        logmel, inst = _chk(logmel, inst)
        b, time_steps, _ = logmel.shape
        wave = torch.empty((b, wave_len), device=logmel.device, dtype=torch.float32)
        scratch = torch.empty((b, time_steps, 3 * bins), device=logmel.device, dtype=torch.float32)
        _lib.call("gs_waveform_generic", _ptr(logmel), _ptr(inst), _ptr(consts["synth_window"]), _ptr(consts["pinv"]), _ptr(wave),
                  _ptr(scratch), b, wave_len, time_steps, int(bins), int(frame_step), _stream())
        return wave

    def pcm16_to_float(self, pcm):
        """Device half of audio_ops.decode_wav (dataset.py:32-36): int16 -> float32 / 32768."""
        if not (torch.is_tensor(pcm) and pcm.is_cuda and pcm.dtype == torch.int16 and pcm.is_contiguous()):
            raise _lib.GansynthLibraryError("pcm16_to_float needs a contiguous CUDA int16 tensor (no CPU path)")
        out = torch.empty(pcm.shape, device=pcm.device, dtype=torch.float32)
        _lib.call("gs_pcm16_to_float", _ptr(pcm), _ptr(out), pcm.numel(), _stream())
        return out

    def waveform_fwd(self, logmel, inst, consts, wave_len, frames_per_segment=None):
        logmel, inst = _chk(logmel, inst)
        b, time_steps, _ = logmel.shape
        if frames_per_segment is None or frames_per_segment >= time_steps:
            frames_per_segment = time_steps
        segs = -(-time_steps // frames_per_segment)
        wave = torch.empty((b, wave_len), device=logmel.device, dtype=torch.float32)
        scratch = torch.empty((b, segs, 1024), device=logmel.device, dtype=torch.float32) if segs > 1 else None
        _lib.call("gs_waveform_fwd", _ptr(logmel), _ptr(inst), _ptr(consts["synth_window"]), _ptr(consts["pb_j0"]),
                  _ptr(consts["pb_cnt"]), _ptr(consts["pb_w"]), int(consts["band"]), _ptr(wave), _ptr(scratch), b,
                  wave_len, time_steps, frames_per_segment, _stream())
        return wave
