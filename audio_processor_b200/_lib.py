"""Loader for libb2a.so — the CUDA (sm_100a) implementation.  There is NO CPU fallback:
if the library is missing or CUDA is unavailable every op raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2a.so")

_lock = threading.Lock()
_lib = None


class B2AError(RuntimeError):
    """Raised for any non-zero return code of the C ABI (message from b2a_last_error())."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libb2a error {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build the sm_100a extension first "
                    "(python -m audio_processor_b200.build, or __graft_entry__.build()). "
                    "audio_processor_b200 has no CPU fallback.")
            _lib = _abi.declare(C.CDLL(LIB_PATH))
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise B2AError(rc, lib().b2a_last_error().decode(errors="replace"))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("audio_processor_b200 needs a CUDA device (B200 / sm_100a); no CPU fallback exists")
    return torch
