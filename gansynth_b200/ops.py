"""Host-side mirror of the reference op library (reference ops.py:149-348) on the CUDA kernels.

Same function names, keyword arguments and defaults as the reference.  Differences, all forced by
leaving TF graph mode:
  * tensors are torch CUDA tensors in NHWC layout ([B, H, W, C]); the reference's ``shape[1]``
    (channels) is ``shape[-1]`` here.  NCHW only exists at the 2-channel image boundary (networks.py);
  * variables live in a VariableStore addressed by the TF variable names the reference's
    tf.variable_scope nesting produces, in the TF layouts ([kh,kw,Cin,Cout], [in,out]);
  * ``get_weight`` returns (variable, runtime multiplier) instead of their product: the equalised-LR
    constant (ops.py:154-160) is folded into the kernels as ``alpha``;
  * conv/dense take an optional fused ``activation`` ("leaky_relu") executed in the kernel epilogue.
The unused normalisers (spectral/weight standardisation, batch/group norm; ops.py:5-146) are out of
scope and rejected.
"""
import contextlib
import math
from collections import OrderedDict

import numpy as np
import torch

from . import functional as F


# ----------------------------------------------------------------------------- variables
class VariableStore(object):
    """name -> leaf tensor, created on first use (tf.get_variable with AUTO_REUSE)."""

    def __init__(self, device="cuda", seed=0):
        self.device = torch.device(device)
        self.vars = OrderedDict()
        self.meta = OrderedDict()     # name -> variance_scale or None (bias)
        self._scope = []
        self._rng = torch.Generator().manual_seed(seed)
        self.flat = {}                # prefix -> flat parameter buffer (after pack())
        self.offsets = {}             # prefix -> name -> (offset, numel) inside the flat buffer

    @contextlib.contextmanager
    def variable_scope(self, name):
        self._scope.append(name)
        try:
            yield
        finally:
            self._scope.pop()

    def scoped(self, name):
        return "/".join(self._scope + [name])

    def get_variable(self, name, shape, initializer):
        full = self.scoped(name)
        var = self.vars.get(full)
        if var is None:
            prefix = full.split("/")[0]
            if prefix in self.flat:
                raise RuntimeError("variable %s requested after scope %r was packed" % (full, prefix))
            host = initializer(tuple(int(s) for s in shape), self._rng)
            var = host.to(self.device, torch.float32).contiguous().requires_grad_(True)
            self.vars[full] = var
        elif tuple(var.shape) != tuple(int(s) for s in shape):
            raise ValueError("variable %s exists with shape %s, requested %s" % (full, tuple(var.shape), tuple(shape)))
        return var

    def trainable_variables(self, scope):
        """tf.get_collection(TRAINABLE_VARIABLES, scope=...) (models.py:78-79)."""
        return OrderedDict((n, v) for n, v in self.vars.items() if n.startswith(scope + "/"))

    def pack(self, scope, alloc=None):
        """Moves every variable of `scope` into one flat fp32 buffer (views keep their names) so the
        optimiser and the data-parallel all-reduce work on a single tensor.  `alloc(numel) -> zeroed fp32 tensor`
        lets the caller place the buffer (e.g. in NVLink symmetric memory)."""
        if scope in self.flat:
            return self.flat[scope]
        items = self.trainable_variables(scope)
        align = 32   # floats: every variable starts on a 128-byte boundary (float4 / TMA friendly)
        offsets, off = OrderedDict(), 0
        for n, v in items.items():
            offsets[n] = (off, v.numel())
            off += -(-v.numel() // align) * align
        flat = alloc(off) if alloc is not None else torch.zeros(off, device=self.device, dtype=torch.float32)
        for n, v in items.items():
            o, k = offsets[n]
            view = flat[o:o + k].view(v.shape)
            view.copy_(v.detach())
            self.vars[n] = view.requires_grad_(True)
        self.flat[scope] = flat
        self.offsets[scope] = offsets
        return flat

    def unflatten(self, scope, flat):
        """name -> view of `flat` (a tensor laid out like the packed parameter buffer of `scope`)."""
        return OrderedDict((n, flat[o:o + k].view(self.vars[n].shape)) for n, (o, k) in self.offsets[scope].items())

    def load(self, values):
        """Injects values (numpy or torch, TF layouts) by variable name."""
        with torch.no_grad():
            for n, val in values.items():
                t = torch.as_tensor(np.asarray(val) if not torch.is_tensor(val) else val)
                self.vars[n].copy_(t.to(self.device, torch.float32))
        F.K.weight_cache_refresh()  # pre-split copies of the old values are stale: re-split them in place

    def state(self):
        return OrderedDict((n, v.detach().cpu()) for n, v in self.vars.items())


_default_store = None


def default_store():
    global _default_store
    if _default_store is None:
        _default_store = VariableStore()
    return _default_store


def set_default_store(store):
    global _default_store
    _default_store = store
    F.K.register_parameters([])     # no parameter buffers are known for the new store yet (drops cached splits)
    return store


def variable_scope(name):
    return default_store().variable_scope(name)


def _truncated_normal(stddev):
    def init(shape, rng):
        t = torch.empty(shape, dtype=torch.float32)
        torch.nn.init.trunc_normal_(t, 0.0, 1.0, -2.0, 2.0, generator=rng)
        return t * stddev
    return init


def _zeros(shape, rng):
    return torch.zeros(shape, dtype=torch.float32)


def _reject_normalisers(apply_weight_standardization, apply_spectral_normalization):
    if apply_spectral_normalization:
        raise NotImplementedError("spectral normalisation (reference ops.py:8-49) is used by no network of the "
                                  "reference (and its body calls tf.indentity, which does not exist)")


def weight_standardization(weight, epsilon=1.0e-12):
    """ops.py:52-66: (w - mean) / sqrt(var + eps) over every axis but the last (tf.nn.moments: biased variance).
    Weight-sized arithmetic of the pitch classifier, done with torch ops."""
    axes = tuple(range(weight.dim() - 1))
    mean = weight.mean(dim=axes, keepdim=True)
    var = weight.var(dim=axes, unbiased=False, keepdim=True)
    return (weight - mean) / torch.sqrt(var + epsilon)


def get_weight(shape, variance_scale=2.0, scale_weight=False, apply_weight_standardization=False,
               apply_spectral_normalization=False):
    """ops.py:149-171.  Returns (variable, alpha): the value the reference would return is
    variable * alpha; alpha is applied inside the kernels.  With weight standardisation (the pitch classifier) the
    returned tensor is the standardised weight."""
    _reject_normalisers(apply_weight_standardization, apply_spectral_normalization)
    stddev = math.sqrt(variance_scale / float(np.prod(shape[:-1])))
    if scale_weight:
        weight, alpha = default_store().get_variable("weight", shape, _truncated_normal(1.0)), stddev
    else:
        weight, alpha = default_store().get_variable("weight", shape, _truncated_normal(stddev)), 1.0
    if apply_weight_standardization:
        weight = weight_standardization(weight).contiguous()
    return weight, alpha


def get_bias(shape):
    """ops.py:174-180."""
    return default_store().get_variable("bias", shape, _zeros)


def _act_code(activation):
    if activation is None:
        return F.ACT_NONE
    if activation in ("leaky_relu", "lrelu"):
        return F.ACT_LRELU
    raise ValueError("unsupported fused activation %r" % (activation,))


# ----------------------------------------------------------------------------- layers
def dense(inputs, units, use_bias=True, variance_scale=2.0, scale_weight=False,
          apply_weight_standardization=False, apply_spectral_normalization=False, activation=None):
    """ops.py:183-201.  inputs [B, in] -> [B, units]."""
    inputs = F.plain(inputs)
    weight, alpha = get_weight([inputs.shape[1], units], variance_scale, scale_weight,
                               apply_weight_standardization, apply_spectral_normalization)
    out = F.DenseF.apply(inputs, weight, alpha)
    act = _act_code(activation)
    if use_bias or act:
        bias = get_bias([units]) if use_bias else None
        out = F.BiasAct.apply(out, bias, act)
    return out


def embedding(inputs, units, variance_scale=2.0, scale_weight=False, apply_weight_standardization=False,
              apply_spectral_normalization=False):
    """ops.py:204-218.  inputs one-hot [B, classes] -> rows of the [classes, units] table."""
    weight, alpha = get_weight([inputs.shape[1], units], variance_scale, scale_weight,
                               apply_weight_standardization, apply_spectral_normalization)
    idx = torch.argmax(inputs, dim=1)
    return F.Embedding.apply(weight, idx, alpha)


def conv2d(inputs, filters, kernel_size, strides=[1, 1], use_bias=True, variance_scale=2.0, scale_weight=False,
           apply_weight_standardization=False, apply_spectral_normalization=False, activation=None,
           pixel_norm_epsilon=None, protocol=False):
    """ops.py:221-247.  NHWC, TF SAME padding, square kernel 1 or 3, stride 1 or 2.  Extensions: `pixel_norm_epsilon`
    appends pixel_normalization (ops.py:330-333) to the fused layer; `protocol` returns the activated output as a
    functional.PreMasked / PnOut wrapper (and `inputs` may be one) so that the NEXT layer's backward applies this
    layer's leaky-relu / pixel-norm gradient in its convolution epilogue -- networks.py uses it internally."""
    ksize, stride = _square(kernel_size), _square(strides)
    weight, alpha = get_weight([ksize, ksize, _channels(inputs), filters], variance_scale, scale_weight,
                               apply_weight_standardization, apply_spectral_normalization)
    bias = get_bias([filters]) if use_bias else None
    return _layer(inputs, weight, bias, "c", ksize, stride, False, alpha, activation, pixel_norm_epsilon, protocol)


def conv2d_transpose(inputs, filters, kernel_size, strides=[1, 1], use_bias=True, variance_scale=2.0,
                     scale_weight=False, apply_weight_standardization=False, apply_spectral_normalization=False,
                     activation=None, pixel_norm_epsilon=None, protocol=False):
    """ops.py:250-280.  The variable is [k, k, Cin, filters] (fan-in from that shape); output is
    [B, H*s, W*s, filters].  `pixel_norm_epsilon`, `protocol`: see conv2d."""
    ksize, stride = _square(kernel_size), _square(strides)
    weight, alpha = get_weight([ksize, ksize, _channels(inputs), filters], variance_scale, scale_weight,
                               apply_weight_standardization, apply_spectral_normalization)
    bias = get_bias([filters]) if use_bias else None
    return _layer(inputs, weight, bias, "t", ksize, stride, True, alpha, activation, pixel_norm_epsilon, protocol)


def _channels(inputs):
    return (inputs.t if isinstance(inputs, (F.PreMasked, F.PnOut)) else inputs).shape[-1]


def _layer(inputs, weight, bias, form, ksize, stride, wswap, alpha, activation, pixel_norm_epsilon, protocol=False):
    act = _act_code(activation)
    if pixel_norm_epsilon is not None:
        if act != F.ACT_LRELU or not F.FUSED_EW:
            return pixel_normalization(F.ConvLayer.apply(F.plain(inputs), weight, bias, form, ksize, stride, wswap, alpha,
                                                         act), pixel_norm_epsilon)
        # conv -> leaky_relu -> pixel_normalization in one kernel; the gradient of the activation pair is applied by
        # whoever consumes the PnOut
        if isinstance(inputs, F.PnOut):
            x, xr = inputs.t, inputs.r
        else:
            x, xr = F.plain(inputs), None
        y, r = F.ConvPnLayer.apply(x, xr, weight, bias, form, ksize, stride, wswap, alpha, pixel_norm_epsilon)
        out = F.PnOut(y, r)
        return out if protocol else F.plain(out)
    x_pm = isinstance(inputs, F.PreMasked) and F.FUSED_EW
    x = inputs.t if x_pm else F.plain(inputs)
    pm_out = bool(protocol) and act == F.ACT_LRELU and F.FUSED_EW
    y = F.ConvLayer.apply(x, weight, bias, form, ksize, stride, wswap, alpha, act, pm_out, x_pm)
    return F.PreMasked(y) if pm_out else y


def _square(v):
    v = [int(a) for a in np.asarray(v).reshape(-1)]
    if len(v) == 1:
        return v[0]
    if len(v) != 2 or v[0] != v[1]:
        raise NotImplementedError("only square kernels / equal strides are supported, got %s" % (v,))
    return v[0]


def upscale2d(inputs, factors=[2, 2]):
    """ops.py:283-291 (NHWC)."""
    inputs = F.plain(inputs)
    factors = np.asanyarray(factors)
    if (factors == 1).all():
        return inputs
    return F.Upscale.apply(inputs, int(factors[0]), int(factors[1]), 1.0)


def downscale2d(inputs, factors=[2, 2]):
    """ops.py:294-305 (NHWC): average pool, kernel = stride = factors."""
    inputs = F.plain(inputs)
    factors = np.asanyarray(factors)
    if (factors == 1).all():
        return inputs
    fh, fw = int(factors[0]), int(factors[1])
    return F.Pool.apply(inputs, fh, fw, 1.0 / (fh * fw))


def pixel_normalization(inputs, epsilon=1.0e-12):
    """ops.py:330-333 over the channel (last) axis."""
    inputs = F.plain(inputs)
    return F.PixelNorm.apply(inputs, epsilon)


def batch_stddev(inputs, groups=4, epsilon=1.0e-12):
    """ops.py:336-348.  inputs [B, H, W, C] -> [B, H, W, 1]; B must be a multiple of `groups`."""
    inputs = F.plain(inputs)
    b = inputs.shape[0]
    if b % groups:
        raise ValueError("batch_stddev: batch %d is not a multiple of groups %d" % (b, groups))
    stat = F.BatchStddev.apply(inputs.reshape(b, -1), groups, epsilon)      # [B/groups]
    per_sample = stat.repeat(groups)                                        # sample n -> stat[n mod B/groups]
    return per_sample.view(b, 1, 1, 1).expand(b, inputs.shape[1], inputs.shape[2], 1)


def leaky_relu(inputs):
    """tf.nn.leaky_relu, alpha 0.2."""
    inputs = F.plain(inputs)
    return F.LeakyRelu.apply(inputs)


def tanh(inputs):
    inputs = F.plain(inputs)
    return F.Tanh.apply(inputs)


# ----------------------------------------------------------------------------- pitch classifier ops
def group_normalization(inputs, groups, epsilon=1.0e-12, relu=False):
    """ops.py:118-146 (NHWC): variables `beta` (zeros) and `gamma` (ones) of shape [C]; `relu` (extension) fuses the
    tf.nn.relu that follows every use in networks.py:318-322, 345-349, 391-396."""
    inputs = F.plain(inputs)
    c = inputs.shape[-1]
    beta = default_store().get_variable("beta", [c], _zeros)
    gamma = default_store().get_variable("gamma", [c], lambda shape, rng: torch.ones(shape, dtype=torch.float32))
    return F.GroupNorm.apply(inputs, gamma, beta, int(groups), float(epsilon), bool(relu))


def max_pooling2d(inputs, kernel_size, strides):
    """ops.py:308-316 (NHWC, TF SAME)."""
    return F.MaxPool.apply(F.plain(inputs), _square(kernel_size), _square(strides))


def reduce_mean_spatial(inputs):
    """tf.reduce_mean(inputs, axis=[2, 3]) of the NCHW reference (networks.py:399): [B, H, W, C] -> [B, C]."""
    return F.SpatialMean.apply(F.plain(inputs))
