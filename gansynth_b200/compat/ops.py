"""Reference module name `ops` -> gansynth_b200.ops (see compat/tensorflow/__init__.py)."""
from gansynth_b200.ops import *  # noqa: F401,F403
from gansynth_b200 import ops as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
