"""CPU oracle for multi-scale deformable attention.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module.  The
product package ``neurips2023_soc_b200`` never does: it raises when its CUDA
library is missing instead of falling back to anything here.

Two restatements of the reference live here:

* ``forward_c`` / ``backward_c`` -- ctypes front end of ``msda_oracle.c``
  (scalar restatement of the reference kernels' per-corner rules,
  /root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:33-159,237-403).
* ``grid_sample_port`` -- the reference's own CPU path restated with
  ``torch.nn.functional.grid_sample``
  (/root/reference/models/ops/functions/ms_deform_attn_func.py:41-61), and
  ``grid_sample_port_grads`` = autograd through it (SURVEY.md 8c "gradient
  oracle").  This is what ``bench.py --impl reference`` times, because the
  reference's Python file itself cannot travel to the GPU box.

Parity pin: the reference ships no golden vectors for this path; both
restatements are pinned against ``tests/golden/*.npz`` which were produced by
importing the reference's ``ms_deform_attn_core_pytorch`` in the build
container (``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

_HERE = Path(__file__).resolve().parent
_SRC = _HERE / "msda_oracle.c"
_LIB = _HERE / "libmsda_oracle.so"
_lib = None


def build_c_oracle(force: bool = False) -> Path:
    """gcc-compile msda_oracle.c into oracle/libmsda_oracle.so (git-ignored)."""
    if force or not _LIB.exists() or _LIB.stat().st_mtime < _SRC.stat().st_mtime:
        cmd = ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", str(_LIB), str(_SRC), "-lm"]
        subprocess.run(cmd, check=True)
    return _LIB


def _load():
    global _lib
    if _lib is None:
        build_c_oracle()
        _lib = ctypes.CDLL(str(_LIB))
        _lib.msda_oracle_threads.restype = ctypes.c_int
    return _lib


def c_oracle_threads() -> int:
    return int(_load().msda_oracle_threads())


def c_oracle_set_threads(n: int) -> None:
    _load().msda_oracle_set_threads(ctypes.c_int(n))


def _np(t, dtype):
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=dtype)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _dims(value, shapes, loc):
    N, S, M, D = value.shape
    L = shapes.shape[0]
    Lq, P = loc.shape[1], loc.shape[4]
    return [ctypes.c_int(int(x)) for x in (N, S, M, D, L, Lq, P)]


def forward_c(value, shapes, lsi, loc, attn, dtype=np.float64):
    """Forward through the C restatement; returns ndarray (N, Lq, M*D)."""
    lib = _load()
    v, lo, at = _np(value, dtype), _np(loc, dtype), _np(attn, dtype)
    sh, ls = _np(shapes, np.int64), _np(lsi, np.int64)
    N, S, M, D = v.shape
    Lq = lo.shape[1]
    out = np.empty((N, Lq, M * D), dtype=dtype)
    fn = lib.msda_oracle_forward_f64 if dtype == np.float64 else lib.msda_oracle_forward_f32
    fn(_ptr(v), _ptr(sh), _ptr(ls), _ptr(lo), _ptr(at), _ptr(out), *_dims(v, sh, lo))
    return out


def backward_c(value, shapes, lsi, loc, attn, grad_out, dtype=np.float64):
    """Backward through the C restatement; returns (grad_value, grad_loc, grad_attn)."""
    lib = _load()
    v, lo, at = _np(value, dtype), _np(loc, dtype), _np(attn, dtype)
    go = _np(grad_out, dtype)
    sh, ls = _np(shapes, np.int64), _np(lsi, np.int64)
    gv, gl, ga = np.empty_like(v), np.empty_like(lo), np.empty_like(at)
    fn = lib.msda_oracle_backward_f64 if dtype == np.float64 else lib.msda_oracle_backward_f32
    fn(_ptr(v), _ptr(sh), _ptr(ls), _ptr(lo), _ptr(at), _ptr(go), _ptr(gv), _ptr(gl), _ptr(ga),
       *_dims(v, sh, lo))
    return gv, gl, ga


def grid_sample_port(value, shapes, loc, attn):
    """The reference's CPU path (ms_deform_attn_func.py:41-61) restated.

    Per level: view that level's rows of ``value`` as an (N*M, D, H, W) image,
    bilinear ``grid_sample`` it (zeros padding, align_corners=False) at
    ``2*loc-1``, then take the attention-weighted sum over the L*P samples.
    """
    N, S, M, D = value.shape
    Lq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    hw = [(int(h), int(w)) for h, w in (shapes.tolist() if hasattr(shapes, "tolist") else shapes)]
    grid = (2.0 * loc - 1.0).permute(0, 2, 1, 3, 4, 5).reshape(N * M, Lq, L, P, 2)
    sampled, start = [], 0
    for lvl, (h, w) in enumerate(hw):
        img = value[:, start:start + h * w].permute(0, 2, 3, 1).reshape(N * M, D, h, w)
        start += h * w
        sampled.append(F.grid_sample(img, grid[:, :, lvl], mode="bilinear",
                                     padding_mode="zeros", align_corners=False))
    sampled = torch.stack(sampled, dim=3)                      # (N*M, D, Lq, L, P)
    wts = attn.permute(0, 2, 1, 3, 4).reshape(N * M, 1, Lq, L, P)
    out = (sampled * wts).sum(dim=(3, 4))                      # (N*M, D, Lq)
    return out.reshape(N, M * D, Lq).transpose(1, 2).contiguous()


def grid_sample_port_grads(value, shapes, loc, attn, grad_out):
    """Autograd through ``grid_sample_port``: (out, grad_value, grad_loc, grad_attn)."""
    v = value.detach().clone().requires_grad_(True)
    lo = loc.detach().clone().requires_grad_(True)
    at = attn.detach().clone().requires_grad_(True)
    out = grid_sample_port(v, shapes, lo, at)
    out.backward(grad_out)
    return out.detach(), v.grad, lo.grad, at.grad


def default_threads() -> int:
    return os.cpu_count() or 1
