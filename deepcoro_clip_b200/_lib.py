"""ctypes binding of libb200clip.so (the C ABI declared in include/b200clip.h).

The library is the product: there is no Python/PyTorch fallback. ``lib()`` raises if the shared object is
missing, and every call raises ``RuntimeError`` on a non-zero status (the reference's convention is Python
exceptions, e.g. utils/registry.py:27-34).
"""
from __future__ import annotations

import ctypes
import re
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
_LIB_PATH = _PKG / "libb200clip.so"
_HEADER = _PKG.parent / "include" / "b200clip.h"
_lib = None
LAUNCHES = 0      # number of b200clip_* kernel-launching calls made by this process (bench.py reports it)
_NO_LAUNCH = {"abi_version", "strerror", "sm_count"}

DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


class B200ClipError(RuntimeError):
    pass


def header_symbols() -> list[str]:
    """Every function name declared in include/b200clip.h (used by the export test)."""
    text = _HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200clip_[a-z0-9_]+)\s*\(", text)))


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise B200ClipError(
                f"{_LIB_PATH} is missing: build it with `python -m deepcoro_clip_b200.build` "
                "(there is deliberately no CPU / PyTorch fallback)")
        _lib = ctypes.CDLL(str(_LIB_PATH))
        _lib.b200clip_strerror.restype = ctypes.c_char_p
        _lib.b200clip_strerror.argtypes = [ctypes.c_int]
    return _lib


def _conv(a):
    if isinstance(a, torch.Tensor):
        return ctypes.c_void_p(a.data_ptr())
    if a is None:
        return ctypes.c_void_p(0)
    if isinstance(a, float):
        return ctypes.c_float(a)
    if isinstance(a, bool):
        return ctypes.c_int(int(a))
    if isinstance(a, int):
        return ctypes.c_int64(a) if abs(a) > 0x7FFFFFFF else ctypes.c_int(a)
    return a


def i64(v: int) -> ctypes.c_int64:
    return ctypes.c_int64(int(v))


def call(name: str, *args) -> None:
    """Calls ``b200clip_<name>`` with tensors converted to device pointers; raises on error."""
    global LAUNCHES
    fn = getattr(lib(), "b200clip_" + name)
    rc = fn(*[_conv(a) for a in args])
    if name not in _NO_LAUNCH:
        LAUNCHES += 1
    if rc != 0:
        msg = lib().b200clip_strerror(rc).decode()
        raise B200ClipError(f"b200clip_{name} failed: {msg} (code {rc})")


def stream_ptr(device=None) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
