"""B200-native NF-DPF particle update (hand-written sm_100a CUDA behind a C-ABI, see include/nfdpf.h)."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
