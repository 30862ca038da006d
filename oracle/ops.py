"""Oracle restatement of the reference's op library (reference ops.py:149-348), PyTorch CPU.

All tensors are NCHW like the reference; weights use the TF variable layouts
([kh, kw, Cin, Cout] for conv, [in, out] for dense) so TF checkpoints could be injected by name.
TF-1.13 semantics follow SURVEY.md Appendix B.  Test infrastructure only (see oracle/__init__.py).
"""
import math

import torch
import torch.nn.functional as F


def he_constant(shape, variance_scale=2.0):
    """Run-time weight multiplier of get_weight(scale_weight=True) (ops.py:154-160):
    sqrt(variance_scale / prod(shape[:-1]))."""
    fan_in = 1
    for s in shape[:-1]:
        fan_in *= int(s)
    return math.sqrt(variance_scale / fan_in)


def same_padding(n, k, s):
    """TF 'SAME' padding (before, after) along one axis."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def dense(x, weight, bias=None, variance_scale=2.0):
    """ops.py:183-201 with scale_weight=True: x @ (W * c) + b."""
    y = x @ (weight * he_constant(weight.shape, variance_scale))
    if bias is not None:
        y = y + bias
    return y


def embedding(onehot, weight, variance_scale=2.0):
    """ops.py:204-218: rows of the scaled table picked by argmax of the one-hot labels; no bias."""
    idx = torch.argmax(onehot, dim=1)
    return (weight * he_constant(weight.shape, variance_scale))[idx]


def conv2d(x, weight, bias=None, strides=(1, 1), variance_scale=2.0):
    """ops.py:221-247: NCHW, SAME, cross-correlation, weight [kh, kw, Cin, Cout] scaled by the He
    constant of its own shape."""
    kh, kw = weight.shape[0], weight.shape[1]
    w = (weight * he_constant(weight.shape, variance_scale)).permute(3, 2, 0, 1)
    pt, pb = same_padding(x.shape[2], kh, strides[0])
    pl, pr = same_padding(x.shape[3], kw, strides[1])
    y = F.conv2d(F.pad(x, (pl, pr, pt, pb)), w, stride=tuple(strides))
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    return y


def conv2d_transpose(x, weight, bias=None, strides=(2, 2), variance_scale=2.0):
    """ops.py:250-280: the variable is [kh, kw, Cin, filters] (fan-in from that shape), transposed
    to [kh, kw, filters, Cin] and used as the filter of tf.nn.conv2d_transpose with output
    [B, filters, H*s, W*s], SAME.  That op is the input-gradient of the SAME strided conv:
    out[n, f, s*i + a - pb, s*j + b - pb] += x[n, c, i, j] * Var[a, b, c, f], cropped to H*s x W*s."""
    kh, kw = weight.shape[0], weight.shape[1]
    w = (weight * he_constant(weight.shape, variance_scale)).permute(2, 3, 0, 1)  # [Cin, f, kh, kw]
    full = F.conv_transpose2d(x, w, stride=tuple(strides))
    oh, ow = x.shape[2] * strides[0], x.shape[3] * strides[1]
    pt, _ = same_padding(oh, kh, strides[0])
    pl, _ = same_padding(ow, kw, strides[1])
    y = full[:, :, pt:pt + oh, pl:pl + ow]
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    return y


def upscale2d(x, factors=(2, 2)):
    """ops.py:283-291: nearest-neighbour repeat; identity when all factors are 1."""
    fh, fw = int(factors[0]), int(factors[1])
    if fh == 1 and fw == 1:
        return x
    return x.repeat_interleave(fh, dim=2).repeat_interleave(fw, dim=3)


def downscale2d(x, factors=(2, 2)):
    """ops.py:294-305: average pool with kernel = stride = factors (sizes divide evenly)."""
    fh, fw = int(factors[0]), int(factors[1])
    if fh == 1 and fw == 1:
        return x
    return F.avg_pool2d(x, kernel_size=(fh, fw), stride=(fh, fw))


def pixel_normalization(x, epsilon=1.0e-12):
    """ops.py:330-333: x / sqrt(mean_c(x^2) + eps) over the channel axis."""
    return x / torch.sqrt(torch.mean(x * x, dim=1, keepdim=True) + epsilon)


def batch_stddev(x, groups=4, epsilon=1.0e-12):
    """ops.py:336-348: reshape to [groups, B/groups, C, H, W]; centre and average squares over the
    groups axis; sqrt(+eps); mean over (C, H, W); tile back to [B, 1, H, W].  Sample n shares its
    statistic with every sample congruent to n modulo B/groups."""
    b, c, h, w = x.shape
    g = x.reshape(groups, -1, c, h, w)
    g = g - g.mean(dim=0, keepdim=True)
    g = (g * g).mean(dim=0)
    g = torch.sqrt(g + epsilon)
    g = g.mean(dim=(1, 2, 3), keepdim=True)
    return g.repeat(groups, 1, h, w)


class MaskTape(object):
    """Parity instrument, not reference behaviour: the leaky-relu masks (x > 0) of a run, in call order.

    `record`: every leaky_relu call appends its mask.  `replay`: every call takes the next mask of the tape
    instead of the sign of its own input -- the run then follows the SAME piecewise-linear branch as the run the
    tape came from (e.g. the CUDA path), which is what separates "a mask flipped because a pre-activation sits
    within rounding of zero" from "the arithmetic differs".  `flips` counts replayed elements whose mask differs
    from this run's own sign, `total` all elements seen."""

    def __init__(self, masks=None):
        self.masks = [] if masks is None else list(masks)
        self.replay = masks is not None
        self.pos = self.flips = self.total = 0

    def __enter__(self):
        global _TAPE
        assert _TAPE is None, "mask tapes do not nest"
        _TAPE = self
        return self

    def __exit__(self, *exc):
        global _TAPE
        _TAPE = None
        if exc[0] is None and self.replay:
            assert self.pos == len(self.masks), "mask tape: %d of %d masks consumed" % (self.pos, len(self.masks))
        return False


_TAPE = None


def leaky_relu(x):
    """tf.nn.leaky_relu default alpha = 0.2."""
    if _TAPE is None:
        return F.leaky_relu(x, 0.2)
    own = x.detach() > 0
    if not _TAPE.replay:
        _TAPE.masks.append(own)
        return F.leaky_relu(x, 0.2)
    assert _TAPE.pos < len(_TAPE.masks), "mask tape exhausted at leaky_relu call %d" % _TAPE.pos
    m = _TAPE.masks[_TAPE.pos]
    _TAPE.pos += 1
    assert m.numel() == x.numel(), "mask tape: call %d has %s, tape has %s" % (_TAPE.pos - 1, tuple(x.shape), tuple(m.shape))
    m = m.reshape(x.shape)
    _TAPE.flips += int((m != own).sum())
    _TAPE.total += x.numel()
    return torch.where(m, x, 0.2 * x)


# ----------------------------------------------------------------------------- pitch classifier ops (ops.py:52-66, 118-146, 308-316)
def weight_standardization(weight, epsilon=1.0e-12):
    """ops.py:52-66: tf.nn.moments over every axis but the last (biased variance)."""
    axes = tuple(range(weight.dim() - 1))
    mean = weight.mean(dim=axes, keepdim=True)
    var = weight.var(dim=axes, unbiased=False, keepdim=True)
    return (weight - mean) / torch.sqrt(var + epsilon)


def conv2d_plain(x, weight, bias=None, strides=(1, 1), standardize=True):
    """ops.py:221-247 with scale_weight=False (the classifier): the variable itself, weight-standardised."""
    kh, kw = weight.shape[0], weight.shape[1]
    w = (weight_standardization(weight) if standardize else weight).permute(3, 2, 0, 1)
    pt, pb = same_padding(x.shape[2], kh, strides[0])
    pl, pr = same_padding(x.shape[3], kw, strides[1])
    y = F.conv2d(F.pad(x, (pl, pr, pt, pb)), w, stride=tuple(strides))
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    return y


def group_normalization(x, gamma, beta, groups, epsilon=1.0e-12):
    """ops.py:118-146 (NCHW): moments over (C/groups, H, W) per (sample, group); gamma / beta are [C]."""
    n, c, h, w = x.shape
    v = x.reshape(n, groups, c // groups, h, w)
    mean = v.mean(dim=(2, 3, 4), keepdim=True)
    var = v.var(dim=(2, 3, 4), unbiased=False, keepdim=True)
    v = ((v - mean) / torch.sqrt(var + epsilon)).reshape(n, c, h, w)
    return v * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)


def max_pooling2d(x, kernel_size, strides):
    """ops.py:308-316: tf.nn.max_pool, SAME (padding never wins)."""
    pt, pb = same_padding(x.shape[2], kernel_size[0], strides[0])
    pl, pr = same_padding(x.shape[3], kernel_size[1], strides[1])
    return F.max_pool2d(F.pad(x, (pl, pr, pt, pb), value=float("-inf")), tuple(kernel_size), tuple(strides))
