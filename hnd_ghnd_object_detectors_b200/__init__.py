"""B200-native GHND hot path.  Host code is Python/PyTorch plumbing; all arithmetic
on the path runs in hand-written sm_100a CUDA behind the C ABI in include/ghnd_b200.h."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
