"""Reference module name `metrics` -> gansynth_b200.metrics (see compat/tensorflow/__init__.py)."""
from gansynth_b200.metrics import *  # noqa: F401,F403
from gansynth_b200 import metrics as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
