"""Reference module name `utils` -> gansynth_b200.utils (see compat/tensorflow/__init__.py)."""
from gansynth_b200.utils import *  # noqa: F401,F403
from gansynth_b200 import utils as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
