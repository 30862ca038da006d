"""Eager restatement of the TensorFlow-1.13 PRIMITIVES that skmhrk1209/GANSynth calls -- TEST INFRASTRUCTURE ONLY.

Purpose: TensorFlow cannot be installed in this image, so the reference cannot run as shipped.  The reference, however, is
nothing but a Python composition of TensorFlow primitives.  With this directory first on sys.path the reference's OWN,
UNMODIFIED ops.py / networks.py / spectral_ops.py / models.py (imported from /root/reference, never copied) execute op by
op on PyTorch-CPU tensors: scopes, variable names and shapes, weight scaling, block order, the tf.cond growth logic, the
loss terms, tf.gradients penalties and the optimizer calls are then the reference's code, not a restatement of it.
tests/golden/make_reference_vectors.py records what that produces; tests/test_reference_pin_cpu.py holds the oracle
(oracle/*.py) to those vectors.  What remains restated -- and is therefore cross-checked separately against
torch / scipy / numpy / torchaudio in tests/test_oracle_cpu.py -- is the behaviour of each primitive below, written from
the TensorFlow-1.13 API documentation (SURVEY.md Appendix B lists the semantics that matter).

Execution model: a graph-mode program is run eagerly.  Building the reference's model object evaluates its graph once on
the tensors its input functions hand out, i.e. ONE construction == ONE session.run; `minimize` returns an object whose
`run()` applies the update computed from that evaluation.  Variables live in a process-wide store keyed by their full
TensorFlow name, so a second construction (tf.AUTO_REUSE) sees the values the first one's train op left behind, exactly
like consecutive session.run calls.  tf.cond evaluates its predicate; with `build_all_branches(True)` it also traces the
untaken branch (discarding the result) the way graph construction does, so that every variable of every branch exists.

Every primitive computes in `float_dtype()` (float32 like the reference, or float64 for sharp comparisons).
Anything the reference does not call raises AttributeError.
"""
import builtins as _b
import contextlib
import math
import re
import types

import numpy as np
import torch
import torch.nn.functional as F

newaxis = None
AUTO_REUSE = "auto_reuse"
float32, float64, int32, int64, string, bool = "float32", "float64", "int32", "int64", "string", "bool"   # noqa: A001

_state = types.SimpleNamespace(dtype=torch.float32, scope=[], variables={}, order=[], both=False, seed=0, draws=0)


def float_dtype():
    return _state.dtype


def set_float_dtype(dtype):
    _state.dtype = dtype


def build_all_branches(flag):
    _state.both = flag


def reset_default_graph():
    _state.scope, _state.variables, _state.order, _state.draws = [], {}, [], 0


def variables():
    """name -> Variable, in creation order (the stand-in's equivalent of tf.global_variables())."""
    return {n: _state.variables[n] for n in _state.order}


# ------------------------------------------------------------------------------------------------ tensors
class Dimension(int):
    @property
    def value(self):
        return int(self)


class TensorShape(tuple):
    def as_list(self):
        return [int(d) for d in self]

    def __getitem__(self, item):
        got = tuple.__getitem__(self, item)
        return TensorShape(got) if isinstance(item, _b.slice) else got

    def concatenate(self, other):
        return TensorShape(tuple(self) + tuple(other))


def _raw(x, like=None):
    """Python / numpy / Tensor -> torch.Tensor, Python floats taking the dtype of the other operand like TF's
    convert_to_tensor does inside a binary op."""
    if isinstance(x, Tensor):
        return x.t
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (_b.bool, np.bool_)):
        return torch.tensor(x)
    if like is not None and (like.is_floating_point() or like.is_complex()):
        return torch.tensor(x, dtype=like.dtype if not like.is_complex() else like.real.dtype)
    if isinstance(x, (int, np.integer)):
        return torch.tensor(int(x), dtype=like.dtype if like is not None else torch.int32)
    if isinstance(x, (float, np.floating)):
        return torch.tensor(float(x), dtype=_state.dtype)
    arr = np.asarray(x)
    return torch.as_tensor(arr, dtype=_state.dtype if arr.dtype.kind == "f" else None)


def _ints(seq):
    return [int(s) for s in seq]


class Tensor(object):
    __array_ufunc__ = None          # numpy scalars defer to the reflected operators below

    def __init__(self, t, name=None):
        self.t = t
        self.name = name

    @property
    def shape(self):
        return TensorShape(Dimension(s) for s in self.t.shape)

    @property
    def dtype(self):
        return str(self.t.dtype).replace("torch.", "")

    def get_shape(self):
        return self.shape

    def set_shape(self, shape):
        want = [None if s is None else int(s) for s in shape]
        assert len(want) == self.t.dim() and all(w is None or w == s for w, s in zip(want, self.t.shape)), (want, self.t.shape)

    def numpy(self):
        return self.t.detach().cpu().numpy()

    def __add__(self, o): return Tensor(self.t + _raw(o, self.t))
    def __radd__(self, o): return Tensor(_raw(o, self.t) + self.t)
    def __sub__(self, o): return Tensor(self.t - _raw(o, self.t))
    def __rsub__(self, o): return Tensor(_raw(o, self.t) - self.t)
    def __mul__(self, o): return Tensor(self.t * _raw(o, self.t))
    def __rmul__(self, o): return Tensor(_raw(o, self.t) * self.t)
    def __truediv__(self, o): return Tensor(self.t / _raw(o, self.t))
    def __rtruediv__(self, o): return Tensor(_raw(o, self.t) / self.t)
    def __neg__(self): return Tensor(-self.t)
    def __getitem__(self, item): return Tensor(self.t[item])

    def __bool__(self):
        raise TypeError("a graph tensor has no Python truth value (tf.cond is the reference's branch)")


class Variable(Tensor):
    def __init__(self, t, name, trainable):
        super().__init__(t, name + ":0")
        self.trainable = trainable
        self.op = types.SimpleNamespace(name=name)

    def assign(self, value):
        with torch.no_grad():
            self.t.copy_(_raw(value, self.t))
        return self


def convert_to_tensor(x, dtype=None):
    return x if isinstance(x, Tensor) else Tensor(_raw(x))


def constant(x, dtype=None):
    return convert_to_tensor(x)


def identity(x, name=None):
    return Tensor(_raw(x), name)


def stop_gradient(x):
    return Tensor(_raw(x).detach())


def cast(x, dtype):
    table = dict(float32=_state.dtype, float64=torch.float64, int32=torch.int32, int64=torch.int64, bool=torch.bool)
    return Tensor(_raw(x).to(table[dtype]))


def divide(x, y):
    """Python-3 true division: integers become float64 first (the growth level of gan_synth_main.py:50-53)."""
    a, b = _raw(x), _raw(y)
    if not a.is_floating_point():
        a = a.double()
    return Tensor(a / (b.to(a.dtype) if isinstance(b, torch.Tensor) else b))


# ------------------------------------------------------------------------------------------------ scopes and variables
@contextlib.contextmanager
def variable_scope(name, reuse=None):
    _state.scope.append(name)
    try:
        yield
    finally:
        _state.scope.pop()


class initializers(object):
    @staticmethod
    def zeros():
        return lambda shape, gen: torch.zeros(shape, dtype=torch.float64)

    @staticmethod
    def ones():
        return lambda shape, gen: torch.ones(shape, dtype=torch.float64)

    @staticmethod
    def truncated_normal(mean=0.0, stddev=1.0):
        """Normal draws further than two standard deviations from the mean are dropped and re-drawn."""
        def init(shape, gen):
            t = torch.empty(shape, dtype=torch.float64)
            torch.nn.init.trunc_normal_(t, float(mean), float(stddev), mean - 2.0 * stddev, mean + 2.0 * stddev, generator=gen)
            return t
        return init

    @staticmethod
    def random_normal(mean=0.0, stddev=1.0):
        return lambda shape, gen: torch.randn(shape, dtype=torch.float64, generator=gen) * stddev + mean


def set_random_seed(seed):
    _state.seed = int(seed)


def get_variable(name, shape=None, initializer=None, trainable=True, dtype=None):
    full = "/".join(_state.scope + [name])
    if full in _state.variables:                         # tf.AUTO_REUSE, the only mode the reference uses
        var = _state.variables[full]
        assert list(var.t.shape) == _ints(shape), (full, var.t.shape, shape)
        return var
    # the draw depends on the graph seed and the variable's NAME, not on creation order
    gen = torch.Generator().manual_seed((_state.seed * 1000003 + sum((i + 1) * b for i, b in enumerate(full.encode()))) % (2 ** 31))
    value = initializer(_ints(shape), gen).to(_state.dtype)
    var = Variable(value.requires_grad_(_b.bool(trainable)), full, _b.bool(trainable))
    _state.variables[full] = var
    _state.order.append(full)
    return var


class GraphKeys(object):
    TRAINABLE_VARIABLES, UPDATE_OPS, TABLE_INITIALIZERS = "trainable_variables", "update_ops", "table_initializer"


def trainable_variables(scope=None):
    return [v for v in variables().values() if v.trainable and (scope is None or re.match(scope, v.op.name))]


def get_collection(key, scope=None):
    if key == GraphKeys.TRAINABLE_VARIABLES:
        return trainable_variables(scope)
    return []


def add_to_collection(key, value):
    pass


@contextlib.contextmanager
def control_dependencies(ops):
    yield


def assign(ref, value):
    return ref.assign(value)


def assign_sub(ref, value):
    return ref.assign(ref.t.detach() - _raw(value, ref.t).detach())


def group(*ops):
    return list(ops)


def local_variables_initializer():
    """Local variables are the two counters of tf.metrics.accuracy."""
    for name in ("accuracy/total", "accuracy/count"):
        _state.variables.pop(name, None)


# ------------------------------------------------------------------------------------------------ control flow
def cond(pred, true_fn, false_fn):
    taken = _b.bool(_raw(pred).item())
    if _state.both:                                      # graph construction traces true_fn, then false_fn
        results = (true_fn(), false_fn())
        return results[0 if taken else 1]
    return true_fn() if taken else false_fn()


def greater(x, y): a = _raw(x); return Tensor(a > _raw(y, a)) if isinstance(x, Tensor) else Tensor(_raw(x, _raw(y)) > _raw(y))
def greater_equal(x, y): a = _raw(x); return Tensor(a >= _raw(y, a))
def less_equal(x, y): a = _raw(x); return Tensor(a <= _raw(y, a))
def equal(x, y): a = _raw(x); return Tensor(a == _raw(y, a))
def logical_and(x, y): return Tensor(_raw(x) & _raw(y))
def reduce_any(x): return Tensor(_raw(x).any())


def where(condition, x=None, y=None):
    c = _raw(condition)
    if x is None:
        return Tensor(torch.nonzero(c))                  # coordinates of the true / non-zero elements, row-major order
    return Tensor(torch.where(c, _raw(x), _raw(y)))


def gather_nd(params, indices):
    p, idx = _raw(params), _raw(indices)
    assert idx.dim() == 2 and idx.shape[1] == p.dim()    # full-rank indices: one scalar per index row
    return Tensor(p[tuple(idx[:, k] for k in range(idx.shape[1]))])


# ------------------------------------------------------------------------------------------------ math
def _unary(fn):
    return lambda x, name=None: Tensor(fn(_raw(x) if isinstance(x, Tensor) else _raw(x).to(_state.dtype)))


log, exp, sqrt, square, sin, cos = (_unary(f) for f in (torch.log, torch.exp, torch.sqrt, torch.square, torch.sin, torch.cos))
abs = _unary(torch.abs)            # noqa: A001  (complex input -> magnitude, like tf.abs)
angle = _unary(torch.angle)
ones_like = _unary(torch.ones_like)


def zeros(shape, dtype=None):
    return Tensor(torch.zeros(_ints(shape), dtype=_state.dtype))


def mod(x, y):
    """FloorMod: the result takes the sign of the divisor."""
    a = _raw(x)
    return Tensor(torch.remainder(a, _raw(y, a)))


def complex(real, imag):           # noqa: A001
    r = _raw(real)
    i = _raw(imag, r)
    return Tensor(torch.complex(r, i.expand_as(r).contiguous()))


def _axes(axis, ndim):
    if axis is None:
        return list(range(ndim))
    return [int(a) % ndim for a in (axis if isinstance(axis, (list, tuple)) else [axis])]


def reduce_mean(x, axis=None, keepdims=False):
    t = _raw(x)
    return Tensor(t.mean(dim=_axes(axis, t.dim()), keepdim=keepdims))


def reduce_sum(x, axis=None, keepdims=False):
    t = _raw(x)
    return Tensor(t.sum(dim=_axes(axis, t.dim()), keepdim=keepdims))


def add_n(xs):
    total = _raw(xs[0])
    for x in xs[1:]:
        total = total + _raw(x)
    return Tensor(total)


def argmax(x, axis=None):
    return Tensor(torch.argmax(_raw(x), dim=int(axis)))


def cumsum(x, axis=0):
    return Tensor(torch.cumsum(_raw(x), dim=int(axis)))


def matmul(a, b, transpose_a=False, transpose_b=False):
    x, y = _raw(a), _raw(b)
    return Tensor((x.t() if transpose_a else x) @ (y.t() if transpose_b else y))


def tensordot(a, b, axes):
    assert axes == 1                                     # the last axis of a against the first axis of b
    x, y = _raw(a), _raw(b)
    return Tensor(torch.tensordot(x, y.to(x.dtype), dims=1))      # float32 constants meet float64 data in float64 mode


def moments(x, axes, keep_dims=False):
    """tf.nn.moments: the variance is the mean squared difference from stop_gradient(mean)."""
    t = _raw(x)
    dims = _axes(axes, t.dim())
    mean = t.mean(dim=dims, keepdim=True)
    var = torch.square(t - mean.detach()).mean(dim=dims, keepdim=True)
    if not keep_dims:
        mean, var = mean.squeeze(dims), var.squeeze(dims)
    return Tensor(mean), Tensor(var)


# ------------------------------------------------------------------------------------------------ shapes
def reshape(tensor, shape):
    return Tensor(_raw(tensor).reshape(_ints(shape)))


def transpose(x, perm):
    return Tensor(_raw(x).permute(*_ints(perm)))


def tile(x, multiples):
    return Tensor(_raw(x).repeat(*_ints(multiples)))


def concat(values, axis):
    return Tensor(torch.cat([_raw(v) for v in values], dim=int(axis)))


def stack(values, axis=0):
    return Tensor(torch.stack([_raw(v) for v in values], dim=int(axis)))


def unstack(value, axis=0):
    return [Tensor(t) for t in torch.unbind(_raw(value), dim=int(axis))]


def squeeze(x, axis=None):
    return Tensor(_raw(x).squeeze() if axis is None else _raw(x).squeeze(int(axis)))


def slice(x, begin, size):         # noqa: A001
    t = _raw(x)
    index = tuple(_b.slice(int(b), None if int(s) == -1 else int(b) + int(s)) for b, s in zip(begin, size))
    return Tensor(t[index])


def pad(x, paddings):
    """Zero padding, one [before, after] pair per axis."""
    t = _raw(x)
    flat = []
    for before, after in reversed([_ints(p) for p in paddings]):
        flat += [before, after]
    if t.is_complex():
        return Tensor(torch.complex(F.pad(t.real, flat), F.pad(t.imag, flat)))
    return Tensor(F.pad(t, flat))


def one_hot(indices, depth):
    return Tensor(F.one_hot(_raw(indices).long(), int(depth)).to(_state.dtype))


def placeholder(dtype, shape, name=None):
    """Stands for an input fed later; traced here with a batch of one."""
    return Tensor(torch.zeros([1 if s is None else int(s) for s in shape], dtype=_state.dtype), name)


# ------------------------------------------------------------------------------------------------ tf.nn
def _same_pads(size, k, s):
    """'SAME': output ceil(size / s); total padding max((out - 1) * s + k - size, 0), the odd element goes AFTER."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def _conv2d_nchw(x, w, strides):
    """Cross-correlation, NCHW input, [kh, kw, in, out] filter, SAME; written from the definition
    out[n, f, i, j] = sum_{a, b, c} xpad[n, c, i * sh + a, j * sw + b] * w[a, b, c, f] over unfolded patches."""
    kh, kw, cin, cout = w.shape
    sh, sw = strides
    n, c, h, wd = x.shape
    assert c == cin
    pt, pb = _same_pads(h, kh, sh)
    pl, pr = _same_pads(wd, kw, sw)
    xp = F.pad(x, (pl, pr, pt, pb))
    oh, ow = -(-h // sh), -(-wd // sw)
    patches = F.unfold(xp, (kh, kw), stride=(sh, sw))                # [n, c * kh * kw, oh * ow], channel-major rows
    patches = patches.reshape(n, c, kh, kw, oh * ow)
    return torch.einsum("ncabp,abcf->nfp", patches, w).reshape(n, cout, oh, ow)


class nn(object):
    @staticmethod
    def conv2d(input, filter, strides, padding, data_format="NHWC"):   # noqa: A002
        assert padding == "SAME" and data_format == "NCHW" and list(strides[:2]) == [1, 1]
        return Tensor(_conv2d_nchw(_raw(input), _raw(filter), _ints(strides[2:])))

    @staticmethod
    def conv2d_transpose(value, filter, output_shape, strides, padding, data_format="NHWC"):   # noqa: A002
        """By definition the gradient of conv2d with respect to its input: `filter` is [kh, kw, out_channels,
        in_channels] and `value` the cotangent of conv2d(input of output_shape, filter, strides, SAME)."""
        assert padding == "SAME" and data_format == "NCHW" and list(strides[:2]) == [1, 1]
        v, w = _raw(value), _raw(filter)
        probe = torch.zeros(_ints(output_shape), dtype=v.dtype, requires_grad=True)
        with torch.enable_grad():
            y = _conv2d_nchw(probe, w, _ints(strides[2:]))
            assert list(y.shape) == list(v.shape), (y.shape, v.shape)
            (dx,) = torch.autograd.grad(y, probe, grad_outputs=v, create_graph=True)
        return Tensor(dx)

    @staticmethod
    def bias_add(value, bias, data_format="NHWC"):
        v, b = _raw(value), _raw(bias)
        if data_format == "NCHW" and v.dim() == 4:
            return Tensor(v + b.reshape(1, -1, 1, 1))
        return Tensor(v + b)

    @staticmethod
    def _pool(value, ksize, strides, padding, data_format, kind):
        assert padding == "SAME" and data_format == "NCHW" and list(ksize[:2]) == [1, 1] and list(strides[:2]) == [1, 1]
        x = _raw(value)
        kh, kw = _ints(ksize[2:])
        sh, sw = _ints(strides[2:])
        pt, pb = _same_pads(x.shape[2], kh, sh)
        pl, pr = _same_pads(x.shape[3], kw, sw)
        if kind == "max":                                 # padding never wins a maximum
            return Tensor(F.max_pool2d(F.pad(x, (pl, pr, pt, pb), value=float("-inf")), (kh, kw), (sh, sw)))
        # average over the elements of the window that lie inside the input
        total = F.avg_pool2d(F.pad(x, (pl, pr, pt, pb)), (kh, kw), (sh, sw), divisor_override=1)
        count = F.avg_pool2d(F.pad(torch.ones_like(x[:1, :1]), (pl, pr, pt, pb)), (kh, kw), (sh, sw), divisor_override=1)
        return Tensor(total / count)

    @staticmethod
    def avg_pool(value, ksize, strides, padding, data_format="NHWC"):
        return nn._pool(value, ksize, strides, padding, data_format, "avg")

    @staticmethod
    def max_pool(value, ksize, strides, padding, data_format="NHWC"):
        return nn._pool(value, ksize, strides, padding, data_format, "max")

    @staticmethod
    def leaky_relu(features, alpha=0.2):
        t = _raw(features)
        return Tensor(torch.where(t > 0, t, t * alpha))   # max(alpha * x, x); derivative alpha at exactly 0

    @staticmethod
    def relu(features):
        return Tensor(torch.relu(_raw(features)))

    @staticmethod
    def tanh(x):
        return Tensor(torch.tanh(_raw(x)))

    @staticmethod
    def softplus(features):
        return Tensor(F.softplus(_raw(features)))

    @staticmethod
    def embedding_lookup(params, ids):
        return Tensor(_raw(params)[_raw(ids)])

    @staticmethod
    def l2_loss(t):
        return Tensor(torch.square(_raw(t)).sum() / 2)

    moments = staticmethod(moments)


class layers(object):
    @staticmethod
    def flatten(inputs):
        t = _raw(inputs)
        return Tensor(t.reshape(t.shape[0], -1))


class losses(object):
    @staticmethod
    def softmax_cross_entropy(onehot_labels, logits):
        """Unit weights, Reduction.SUM_BY_NONZERO_WEIGHTS: the mean over the batch."""
        per_example = -(torch.log_softmax(_raw(logits), dim=-1) * _raw(onehot_labels)).sum(dim=-1)
        return Tensor(per_example.mean())


class metrics(object):
    @staticmethod
    def accuracy(labels, predictions):
        """Streaming: (value before this batch, update_op = value once this batch is counted), over two local variables."""
        total = _state.variables.setdefault("accuracy/total", Variable(torch.zeros((), dtype=torch.float64), "accuracy/total", False))
        count = _state.variables.setdefault("accuracy/count", Variable(torch.zeros((), dtype=torch.float64), "accuracy/count", False))
        before = total.t / count.t if float(count.t) else torch.zeros((), dtype=torch.float64)
        hits = (_raw(labels) == _raw(predictions)).double()
        total.t, count.t = total.t + hits.sum(), count.t + hits.numel()
        return Tensor(before.to(_state.dtype)), Tensor((total.t / count.t).to(_state.dtype))


def gradients(ys, xs):
    """Sum of d(sum of every y)/dx for each x, differentiable again."""
    ys = [_raw(y) for y in (ys if isinstance(ys, (list, tuple)) else [ys])]
    total = sum(y.sum() for y in ys)
    got = torch.autograd.grad(total, [_raw(x) for x in xs], create_graph=True, allow_unused=True)
    return [None if g is None else Tensor(g) for g in got]


class random(object):
    @staticmethod
    def normal(shape, mean=0.0, stddev=1.0):
        _state.draws += 1
        gen = torch.Generator().manual_seed(_state.seed * 7919 + _state.draws)
        t = torch.randn(_ints(shape), dtype=torch.float64, generator=gen) * stddev + mean
        return Tensor(t.to(_state.dtype).requires_grad_(True))


# ------------------------------------------------------------------------------------------------ tf.signal
class signal(object):
    @staticmethod
    def hann_window(window_length, periodic=True, dtype=None):
        n = torch.arange(int(window_length), dtype=torch.float64)
        denom = window_length if periodic else window_length - 1
        return Tensor((0.5 - 0.5 * torch.cos(2.0 * math.pi * n / denom)).to(_state.dtype))

    @staticmethod
    def _fft_length(frame_length):
        return 1 << int(math.ceil(math.log2(frame_length)))           # smallest power of two enclosing the frame

    @staticmethod
    def stft(signals, frame_length, frame_step, fft_length=None, window_fn=None, pad_end=False):
        """frames of frame_length every frame_step (no end padding), times the window, rfft of fft_length."""
        assert not pad_end
        x = _raw(signals)
        fft_length = fft_length or signal._fft_length(frame_length)
        frames = x.unfold(-1, int(frame_length), int(frame_step))
        if window_fn is not None:
            frames = frames * _raw(window_fn(frame_length, dtype=None))
        return Tensor(torch.fft.rfft(frames, n=fft_length, dim=-1))

    @staticmethod
    def inverse_stft_window_fn(frame_step, forward_window_fn):
        """forward window / (sum over the overlapping hops of its square), per position inside a hop."""
        def window_fn(frame_length, dtype=None):
            fw = _raw(forward_window_fn(frame_length, dtype=dtype))
            overlaps = -(-frame_length // frame_step)
            denom = F.pad(torch.square(fw), (0, overlaps * frame_step - frame_length))
            denom = denom.reshape(overlaps, frame_step).sum(dim=0, keepdim=True).repeat(overlaps, 1).reshape(-1)
            return Tensor(fw / denom[:frame_length])
        return window_fn

    @staticmethod
    def inverse_stft(stfts, frame_length, frame_step, fft_length=None, window_fn=None):
        """irfft of fft_length cut to frame_length, times the window, overlap-added every frame_step."""
        s = _raw(stfts)
        fft_length = fft_length or signal._fft_length(frame_length)
        frames = torch.fft.irfft(s, n=fft_length, dim=-1)[..., :frame_length]
        if window_fn is not None:
            frames = frames * _raw(window_fn(frame_length, dtype=None))
        count = frames.shape[-2]
        out = torch.zeros(list(frames.shape[:-2]) + [frame_step * (count - 1) + frame_length], dtype=frames.dtype)
        for k in range(count):
            out[..., k * frame_step:k * frame_step + frame_length] += frames[..., k, :]
        return Tensor(out)

    @staticmethod
    def linear_to_mel_weight_matrix(num_mel_bins, num_spectrogram_bins, sample_rate, lower_edge_hertz, upper_edge_hertz,
                                    dtype="float32"):
        """HTK mel scale 1127 ln(1 + f / 700); bin 0 (DC) gets zero weight; triangles are drawn in MEL space between
        num_mel_bins + 2 equally spaced edges.  TF 1.13 evaluates every step in `dtype` -- float32 unless the caller says
        otherwise, and the reference does not -- so the result is a float32 constant whatever float_dtype() is."""
        assert dtype == "float32"
        f32 = np.float32

        def hertz_to_mel(f):
            return (f32(1127.0) * np.log(f32(1.0) + f / f32(700.0))).astype(f32)
        nyquist = f32(sample_rate / 2.0)
        linear = np.linspace(f32(0.0), nyquist, num_spectrogram_bins, dtype=f32)[1:]
        spec_mel = hertz_to_mel(linear)[:, None]
        edges = np.linspace(hertz_to_mel(np.asarray(lower_edge_hertz, f32)), hertz_to_mel(np.asarray(upper_edge_hertz, f32)),
                            num_mel_bins + 2, dtype=f32)
        lower, center, upper = edges[:-2][None, :], edges[1:-1][None, :], edges[2:][None, :]
        up = (spec_mel - lower) / (center - lower)
        down = (upper - spec_mel) / (upper - center)
        weights = np.maximum(f32(0.0), np.minimum(up, down)).astype(f32)
        return Tensor(torch.from_numpy(np.pad(weights, [[1, 0], [0, 0]])))


# ------------------------------------------------------------------------------------------------ tf.train
class _TrainOp(object):
    def __init__(self, apply):
        self.run = apply


class _Optimizer(object):
    def _slot(self, var, suffix, init=0.0):
        name = var.op.name + "/" + suffix
        if name not in _state.variables:
            _state.variables[name] = Variable(torch.full_like(var.t.detach(), init), name, False)
            _state.order.append(name)
        return _state.variables[name]

    def _scalar(self, name, init):
        if name not in _state.variables:
            _state.variables[name] = Variable(torch.tensor(init, dtype=torch.float64), name, False)
            _state.order.append(name)
        return _state.variables[name]

    def minimize(self, loss, var_list=None, global_step=None):
        var_list = list(var_list) if var_list is not None else trainable_variables()
        grads = torch.autograd.grad(_raw(loss), [v.t for v in var_list], allow_unused=True, retain_graph=True)
        pairs = [(None if g is None else g.detach().clone(), v) for g, v in zip(grads, var_list)]
        op = _TrainOp(lambda: self._apply(pairs, global_step))
        op.grads_and_vars = pairs
        return op


class train(object):
    @staticmethod
    def get_or_create_global_step():
        if "global_step" not in _state.variables:
            _state.variables["global_step"] = Variable(torch.zeros((), dtype=torch.int64), "global_step", False)
            _state.order.append("global_step")
        return _state.variables["global_step"]

    create_global_step = get_global_step = get_or_create_global_step

    @staticmethod
    def exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False):
        p = _raw(global_step).double() / float(decay_steps)
        return Tensor(learning_rate * decay_rate ** (torch.floor(p) if staircase else p))

    class AdamOptimizer(_Optimizer):
        """m, v slots per variable, beta powers per optimizer; lr_t = lr sqrt(1 - b2^t) / (1 - b1^t);
        var -= lr_t m / (sqrt(v) + epsilon): epsilon sits OUTSIDE the bias correction."""

        def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
            self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon

        def _apply(self, pairs, global_step):
            first = pairs[0][1].op.name
            p1, p2 = self._scalar(first + "/beta1_power", self.b1), self._scalar(first + "/beta2_power", self.b2)
            lr = float(_raw(self.lr))
            lr_t = lr * math.sqrt(1.0 - float(p2.t)) / (1.0 - float(p1.t))
            with torch.no_grad():
                for g, var in pairs:
                    if g is None:
                        continue
                    m, v = self._slot(var, "Adam"), self._slot(var, "Adam_1")
                    m.t.mul_(self.b1).add_(g * (1.0 - self.b1))
                    v.t.mul_(self.b2).add_(g * g * (1.0 - self.b2))
                    var.t.sub_(lr_t * m.t / (torch.sqrt(v.t) + self.eps))
                p1.t.mul_(self.b1)
                p2.t.mul_(self.b2)
                if global_step is not None:
                    global_step.t.add_(1)

    class MomentumOptimizer(_Optimizer):
        """accum = momentum * accum + g;  var -= lr * accum, or with use_nesterov  var -= lr * g + lr * momentum * accum."""

        def __init__(self, learning_rate, momentum, use_nesterov=False):
            self.lr, self.momentum, self.nesterov = learning_rate, momentum, use_nesterov

        def _apply(self, pairs, global_step):
            lr = float(_raw(self.lr))
            with torch.no_grad():
                for g, var in pairs:
                    if g is None:
                        continue
                    accum = self._slot(var, "Momentum")
                    accum.t.mul_(self.momentum).add_(g)
                    var.t.sub_(lr * g + lr * self.momentum * accum.t if self.nesterov else lr * accum.t)
                if global_step is not None:
                    global_step.t.add_(1)
