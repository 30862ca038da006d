"""Reference module name `spectral_ops` -> gansynth_b200.spectral_ops (see compat/tensorflow/__init__.py)."""
from gansynth_b200.spectral_ops import *  # noqa: F401,F403
from gansynth_b200 import spectral_ops as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
