"""ctypes binding of libb200clip.so (the C ABI declared in include/b200clip.h).

The library is the product: there is no Python/PyTorch fallback. ``lib()`` raises if the shared object is
missing, and every call raises ``RuntimeError`` on a non-zero status (the reference's convention is Python
exceptions, e.g. utils/registry.py:27-34).
"""
from __future__ import annotations

import ctypes
import os
import re
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
_LIB_PATH = _PKG / "libb200clip.so"
_HEADER = _PKG.parent / "include" / "b200clip.h"
_lib = None
LAUNCHES = 0      # number of b200clip_* kernel-launching calls made by this process (bench.py reports it)
_NO_LAUNCH = {"abi_version", "strerror", "sm_count", "attnpool_splits", "attnpool_bwd_splits", "retrieval_segments",
              "rowlse_slots", "milpool_plan", "milpool_tc_plan", "gstore_elems", "aggregator_sizes"}

DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


class B200ClipError(RuntimeError):
    pass


def header_symbols() -> list[str]:
    """Every function name declared in include/b200clip.h (used by the export test)."""
    text = _HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200clip_[a-z0-9_]+)\s*\(", text)))


_CTYPE = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "float": ctypes.c_float, "double": ctypes.c_double}


def _prototypes() -> dict:
    """{name: (restype, [argtypes])} parsed from include/b200clip.h, so ctypes converts every argument in C (and a
    Python int can never be silently truncated into the wrong width)."""
    text = re.sub(r"/\*.*?\*/", "", _HEADER.read_text(), flags=re.S)
    out = {}
    for ret, name, args in re.findall(r"\b(int|const char\*)\s+(b200clip_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        types = []
        for a in [x.strip() for x in args.replace("\n", " ").split(",")]:
            if a == "void":
                continue
            t = re.sub(r"\s+[A-Za-z_0-9]+$", "", a).replace("const ", "").strip()
            types.append(ctypes.c_void_p if t.endswith("*") else _CTYPE[t])
        out[name] = (ctypes.c_char_p if ret != "int" else ctypes.c_int, types)
    return out


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise B200ClipError(
                f"{_LIB_PATH} is missing: build it with `python -m deepcoro_clip_b200.build` "
                "(there is deliberately no CPU / PyTorch fallback)")
        _lib = ctypes.CDLL(str(_LIB_PATH))
        for name, (ret, types) in _prototypes().items():
            fn = getattr(_lib, name, None)
            if fn is not None:
                fn.restype = ret
                fn.argtypes = types
    return _lib


_FN: dict = {}
_Tensor = torch.Tensor


def i64(v: int) -> int:
    return int(v)


# B200CLIP_NVTX=1: every ABI call is wrapped in an NVTX range named after the entry point, so nsys / ncu timelines show
# the library's kernels grouped by the call that launched them (the reference has no tracing hooks at all).
_NVTX = os.environ.get("B200CLIP_NVTX", "0") == "1"


def call(name: str, *args) -> None:
    """Calls ``b200clip_<name>``: tensors become device pointers, everything else is converted by ctypes against the
    header prototype; raises on a non-zero status."""
    global LAUNCHES
    fn = _FN.get(name)
    if fn is None:
        fn = _FN[name] = getattr(lib(), "b200clip_" + name)
    if _NVTX:
        torch.cuda.nvtx.range_push("b200clip_" + name)
        try:
            rc = fn(*[a.data_ptr() if type(a) is _Tensor or isinstance(a, _Tensor) else a for a in args])
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        rc = fn(*[a.data_ptr() if isinstance(a, _Tensor) else a for a in args])
    if name not in _NO_LAUNCH:
        LAUNCHES += 1
    if rc != 0:
        msg = lib().b200clip_strerror(rc).decode()
        raise B200ClipError(f"b200clip_{name} failed: {msg} (code {rc})")


def try_call(name: str, *args) -> bool:
    """Like ``call`` for entry points that answer B2_ENOSYS (-38) when a shape does not qualify for a specialised kernel:
    returns False in that case (nothing was launched) so the caller can take the general path; raises on any other error."""
    global LAUNCHES
    fn = _FN.get(name)
    if fn is None:
        fn = _FN[name] = getattr(lib(), "b200clip_" + name)
    rc = fn(*[a.data_ptr() if isinstance(a, _Tensor) else a for a in args])
    if rc == -38:
        return False
    if name not in _NO_LAUNCH:
        LAUNCHES += 1
    if rc != 0:
        raise B200ClipError(f"b200clip_{name} failed: {lib().b200clip_strerror(rc).decode()} (code {rc})")
    return True


def stream_ptr(device=None) -> int:
    """Raw cudaStream_t of torch's current stream on ``device``."""
    idx = device.index if isinstance(device, torch.device) and device.index is not None else (
        device if isinstance(device, int) else torch.cuda.current_device())
    return torch._C._cuda_getCurrentRawStream(idx)
