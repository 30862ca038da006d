"""Reference module name `networks` -> gansynth_b200.networks (see compat/tensorflow/__init__.py)."""
from gansynth_b200.networks import *  # noqa: F401,F403
from gansynth_b200 import networks as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
