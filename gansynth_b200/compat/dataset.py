"""Reference module name `dataset` -> gansynth_b200.dataset (see compat/tensorflow/__init__.py)."""
from gansynth_b200.dataset import *  # noqa: F401,F403
from gansynth_b200 import dataset as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
