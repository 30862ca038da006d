"""A `tensorflow` stand-in with exactly the names the reference `gan_synth_main.py` touches (SURVEY App. C;
gan_synth_main.py:39-54, 69, 91-98, 113-114) -- and `pitch_classifier_main.py` (:33-37, 69-74, 80-87) -- so that those files
run UNCHANGED on gansynth_b200:

    PYTHONPATH=<repo>/gansynth_b200/compat:<repo> python /path/to/reference/gan_synth_main.py --train ...

`compat/` also holds `dataset.py`, `models.py`, `networks.py`, `utils.py`, `ops.py`, `spectral_ops.py`, `metrics.py`:
the reference's top-level module names re-exporting this package.  Nothing here computes: the graph / session
scaffolding of TF-1 has no counterpart (the model runs eagerly or as CUDA graphs), the global step is the package's
own counter, `growing_level` stays a lazy scalar that is re-read at every forward pass, and `tf.random.normal` inside
`fake_input_fn`'s lambda draws a fresh batch at every call.  This is NOT TensorFlow: any other attribute raises.
"""
import contextlib as _contextlib
import types as _types

import torch as _torch

from gansynth_b200 import models as _models

__version__ = "0.0+gansynth_b200.compat"
float32 = _torch.float32
int32 = _torch.int32
int64 = _torch.int64

logging = _types.SimpleNamespace(INFO=20, set_verbosity=lambda level: None)


class Graph(object):
    """tf.Graph(): `with tf.Graph().as_default():` scopes one model build; here it resets the package-level state a
    fresh graph would not see (global step, cached input pipelines)."""

    @_contextlib.contextmanager
    def as_default(self):
        from gansynth_b200 import dataset
        _models.reset_global_step()
        dataset.reset_pipelines()
        yield self


def set_random_seed(seed):
    _torch.manual_seed(int(seed))


def cast(x, dtype):
    """Only used as tf.cast(global_step / growing_steps, tf.float32): the lazy scalar passes through."""
    return x


def divide(x, y):
    return x / y


class _Train(object):
    @staticmethod
    def create_global_step():
        return _models.get_or_create_global_step()

    get_or_create_global_step = create_global_step
    get_global_step = create_global_step

    @staticmethod
    def exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False, name=None):
        """pitch_classifier_main.py:69-74 (called inside hyper_params.learning_rate's lambda with the live global step)."""
        return _models.exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase)


train = _Train()


class _Random(object):
    @staticmethod
    def normal(shape, mean=0.0, stddev=1.0):
        device = "cuda" if _torch.cuda.is_available() else "cpu"
        return _torch.randn(*[int(s) for s in shape], device=device) * stddev + mean


random = _Random()


class _Options(object):
    """tf.ConfigProto / tf.GPUOptions: an inert bag of keyword options (gan_synth_main.py:91-98)."""

    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)


ConfigProto = _Options
GPUOptions = _Options


class GraphDef(object):
    """tf.GraphDef.FromString(bytes) (gan_synth_main.py:113-114): kept as opaque bytes.  GANSynth.evaluate needs the
    classifier as a callable and says so when handed one of these."""

    def __init__(self, data=b""):
        self.data = data

    @classmethod
    def FromString(cls, data):
        return cls(data)


def __getattr__(name):
    # TensorBoard's compat layer imports `tensorflow` when it can and expects tf.io.gfile & co.; it ships a stub of
    # exactly those pieces for TensorFlow-less installs -- hand it its own stub so that SummaryWriter keeps working
    # while this stand-in is on sys.path
    try:
        from tensorboard.compat import tensorflow_stub as _stub
        if hasattr(_stub, name):
            return getattr(_stub, name)
    except Exception:
        pass
    raise AttributeError("gansynth_b200.compat.tensorflow is a stand-in for the names gan_synth_main.py uses; "
                         "`tf.%s` is not one of them" % name)
