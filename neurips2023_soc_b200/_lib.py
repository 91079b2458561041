"""ctypes binding of libmsda_b200.so (include/msda_b200.h).  No fallback of any kind:
if the library is missing or a call fails, this module raises."""
from __future__ import annotations

import ctypes
from pathlib import Path

import os

PKG = Path(__file__).resolve().parent
# MSDA_LIB selects another build of the same library (kernel-variant A/B runs in tools/)
LIB_PATH = Path(os.environ["MSDA_LIB"]) if os.environ.get("MSDA_LIB") else PKG / "libmsda_b200.so"

F32, BF16, F16, F64 = 0, 1, 2, 3
FLAG_PYRAMID_TILES, FLAG_GENERIC, FLAG_ATOMIC_GRAD_VALUE, FLAG_BF16_VEC4 = 1, 2, 4, 8
FLAG_WALK_DENSE = 16
FLAG_BIN_KERNEL = 32
FLAG_DIRECT_SPLIT = 64
FLAG_UNORDERED = 128

EXPORTS = (
    "msda_version", "msda_last_error", "msda_forward", "msda_forward_ex",
    "msda_backward_workspace_bytes", "msda_backward", "msda_backward_ex", "msda_last_launch_count",
    "msda_profile_enable", "msda_profile_read",
    "msda_index_bytes", "msda_forward_indexed", "msda_backward_indexed", "msda_forward_fused", "msda_backward_fused",
    "msda_backward_fused_raw", "msda_add_layernorm_forward", "msda_add_layernorm_backward", "msda_add_layernorm_backward_workspace_bytes",
)

_lib = None


class MSDAError(RuntimeError):
    """A non-zero status from the C ABI (the reference raises RuntimeError through AT_ASSERTM)."""


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise MSDAError(
            f"{LIB_PATH} is missing. Build it with `python -m neurips2023_soc_b200.build` "
            "(or __graft_entry__.build()). There is no CPU or PyTorch fallback for this op.")
    lib = ctypes.CDLL(str(LIB_PATH))
    vp, i, u, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_size_t
    dims = [i] * 7
    lib.msda_version.restype = i
    lib.msda_last_error.restype = ctypes.c_char_p
    lib.msda_last_launch_count.restype = i
    lib.msda_forward.restype = i
    lib.msda_forward.argtypes = [vp] * 6 + dims + [i, i, i, vp]
    lib.msda_forward_ex.restype = i
    lib.msda_forward_ex.argtypes = [vp] * 6 + dims + [i, i, i, vp, u]
    lib.msda_backward_workspace_bytes.restype = sz
    lib.msda_backward_workspace_bytes.argtypes = dims + [i, i]
    lib.msda_backward.restype = i
    lib.msda_backward.argtypes = [vp] * 10 + [sz] + dims + [i, i, i, vp]
    lib.msda_backward_ex.restype = i
    lib.msda_backward_ex.argtypes = [vp] * 10 + [sz] + dims + [i, i, i, vp, u]
    lib.msda_index_bytes.restype = sz
    lib.msda_index_bytes.argtypes = dims
    lib.msda_forward_indexed.restype = i
    lib.msda_forward_indexed.argtypes = [vp] * 7 + [sz] + dims + [i, i, i, vp, u]
    lib.msda_forward_fused.restype = i
    lib.msda_forward_fused.argtypes = [vp] * 11 + [sz] + dims + [i, i, i, vp, u]
    lib.msda_backward_indexed.restype = i
    lib.msda_backward_indexed.argtypes = [vp] * 10 + [sz, vp, sz] + dims + [i, i, i, vp, u]
    lib.msda_backward_fused.restype = i
    lib.msda_backward_fused.argtypes = [vp] * 11 + [sz, vp, sz] + dims + [i, i, i, vp, u]
    lib.msda_backward_fused_raw.restype = i
    lib.msda_backward_fused_raw.argtypes = [vp] * 12 + [sz, vp, sz] + dims + [i, i, i, vp, u]
    ll, fl, fp = ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p
    lib.msda_add_layernorm_forward.restype = i
    lib.msda_add_layernorm_forward.argtypes = [vp] * 8 + [ll, i, i, fl, vp]
    lib.msda_add_layernorm_backward_workspace_bytes.restype = sz
    lib.msda_add_layernorm_backward_workspace_bytes.argtypes = [ll, i]
    lib.msda_add_layernorm_backward.restype = i
    lib.msda_add_layernorm_backward.argtypes = [vp] * 9 + [sz, ll, i, i, fp]
    lib.msda_profile_enable.restype = None
    lib.msda_profile_enable.argtypes = [i]
    lib.msda_profile_read.restype = i
    lib.msda_profile_read.argtypes = [ctypes.c_char_p, sz, ctypes.POINTER(ctypes.c_float), i]
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().msda_last_error().decode("utf-8", "replace")
        raise MSDAError(f"libmsda_b200 status {status}: {msg}")


def profile_enable(on: bool) -> None:
    load().msda_profile_enable(1 if on else 0)


def profile_read(cap: int = 4096):
    """[(kernel name, milliseconds)] recorded since profile_enable(True)."""
    names = ctypes.create_string_buffer(64 * cap)
    ms = (ctypes.c_float * cap)()
    n = load().msda_profile_read(names, len(names), ms, cap)
    labels = names.value.decode().split("\n")[:n]
    return list(zip(labels, [float(ms[k]) for k in range(n)]))
