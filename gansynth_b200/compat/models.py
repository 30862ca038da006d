"""Reference module name `models` -> gansynth_b200.models (see compat/tensorflow/__init__.py)."""
from gansynth_b200.models import *  # noqa: F401,F403
from gansynth_b200 import models as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
