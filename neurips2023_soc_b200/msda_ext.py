"""Host side of the operator boundary: the two functions the reference binds from its
compiled extension (/root/reference/models/ops/src/vision.cpp:13-16), with the same
names, argument order and error behaviour, routed to libmsda_b200.so through ctypes.

    ms_deform_attn_forward(value, spatial_shapes, level_start_index,
                           sampling_loc, attn_weight, im2col_step) -> Tensor (N, Lq, M*D)
    ms_deform_attn_backward(value, spatial_shapes, level_start_index,
                            sampling_loc, attn_weight, grad_output, im2col_step)
                           -> [grad_value, grad_sampling_loc, grad_attn_weight]

Host logic mirrored from ms_deform_attn_cuda_forward / _backward
(/root/reference/models/ops/src/cuda/ms_deform_attn_cuda.cu:20-80, :83-153):
contiguity and device checks (:28-38, :93-105) raise RuntimeError, CPU tensors raise
"Not implemented on the CPU" (/root/reference/models/ops/src/ms_deform_attn.h:38,60), the
callee allocates the results with the dtype/device of ``value`` (:54, :121-123) and the
work is queued on the current CUDA stream (:65, :135) without synchronising.

torch is used for device memory and the stream handle only.
"""
from __future__ import annotations

from typing import List

import torch

from . import _lib

_DTYPE = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16, torch.float64: _lib.F64}

#: bench/diagnostic switch: flags forwarded to msda_*_ex (0 in production)
DEFAULT_FLAGS = 0


def _require(cond: bool, msg: str) -> None:
    if not cond:
        raise RuntimeError(msg)


def _check_inputs(named):
    for name, t in named:
        _require(isinstance(t, torch.Tensor), f"{name} must be a tensor")
        _require(t.is_contiguous(), f"{name} tensor has to be contiguous")
    if not named[0][1].is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    dev = named[0][1].device
    for name, t in named:
        _require(t.is_cuda, f"{name} must be a CUDA tensor")
        _require(t.device == dev, f"{name} must be on {dev}, found {t.device}")


def _problem(value, spatial_shapes, level_start_index, sampling_loc, attn_weight):
    _require(value.dim() == 4, "value must be (N, S, M, D)")
    _require(sampling_loc.dim() == 6 and sampling_loc.shape[-1] == 2, "sampling_loc must be (N, Lq, M, L, P, 2)")
    _require(attn_weight.dim() == 5, "attn_weight must be (N, Lq, M, L, P)")
    _require(spatial_shapes.dtype == torch.int64 and level_start_index.dtype == torch.int64,
             "spatial_shapes and level_start_index must be int64")
    N, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    _, Lq, M2, L2, P, _ = sampling_loc.shape
    _require(spatial_shapes.dim() == 2 and spatial_shapes.shape[1] == 2, "spatial_shapes must be (L, 2)")
    _require(level_start_index.numel() == L, "level_start_index must have L entries")
    _require(sampling_loc.shape[0] == N and M2 == M and L2 == L, "sampling_loc does not match value / spatial_shapes")
    _require(tuple(attn_weight.shape) == (N, Lq, M, L, P), "attn_weight does not match sampling_loc")
    _require(value.dtype in _DTYPE, f"unsupported value dtype {value.dtype}")
    _require(sampling_loc.dtype == attn_weight.dtype, "sampling_loc and attn_weight must share a dtype")
    _require(sampling_loc.dtype == value.dtype or (sampling_loc.dtype == torch.float32 and value.dtype != torch.float64),
             f"sampling_loc/attn_weight must be {value.dtype} or float32, found {sampling_loc.dtype}")
    return (N, S, M, D, L, Lq, P), _DTYPE[value.dtype], _DTYPE[sampling_loc.dtype]


def _into(buf, shape, dtype, device, what):
    """A caller-owned result buffer (``out=`` style, beyond the reference signature) viewed with the result's shape."""
    n = 1
    for s in shape:
        n *= int(s)
    _require(isinstance(buf, torch.Tensor) and buf.is_cuda and buf.device == device and buf.dtype == dtype
             and buf.is_contiguous() and buf.numel() == n,
             f"{what} must be a contiguous {dtype} tensor of {n} elements on {device}")
    return buf.view(shape)


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                           im2col_step: int, flags: int | None = None, want_index: bool = False,
                           out=None, index_buf=None):
    """``want_index=True`` (not part of the reference signature) also returns the index the matching
    backward can use -- and uses up: ``(output, index)``, see msda_forward_indexed in include/msda_b200.h.
    ``out`` / ``index_buf``: caller-owned buffers to write into instead of allocating (callers that run the op
    in a loop on their own streams, e.g. host_frames.HostFramePipeline)."""
    _check_inputs([("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                   ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)])
    dims, vdt, adt = _problem(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    N, S, M, D, L, Lq, P = dims
    lib = _lib.load()
    with torch.cuda.device(value.device):
        out = (torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device) if out is None
               else _into(out, (N, Lq, M * D), value.dtype, value.device, "out"))
        stream = torch.cuda.current_stream().cuda_stream
        index, index_ptr, index_bytes = None, None, 0
        if want_index:
            index_bytes = int(lib.msda_index_bytes(*dims))     # 0: calls this small keep no index (direct gather)
            if index_bytes and index_buf is None:
                index = torch.empty(index_bytes, dtype=torch.uint8, device=value.device)
            elif index_bytes:
                _require(index_buf.is_cuda and index_buf.dtype == torch.uint8 and index_buf.is_contiguous()
                         and index_buf.numel() >= index_bytes, f"index_buf must hold {index_bytes} bytes")
                index = index_buf[:index_bytes]
            if index is not None:
                index_ptr = index.data_ptr()
        _lib.check(lib.msda_forward_indexed(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            sampling_loc.data_ptr(), attn_weight.data_ptr(), out.data_ptr(), index_ptr, index_bytes,
            *dims, vdt, adt, int(im2col_step), stream, DEFAULT_FLAGS if flags is None else flags))
    return (out, index) if want_index else out


def backward_workspace_bytes(value, sampling_loc) -> int:
    """msda_backward_workspace_bytes for a call with these tensors (for callers that own the workspace)."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_loc.shape
    return int(_lib.load().msda_backward_workspace_bytes(N, S, M, D, L, Lq, P, _DTYPE[value.dtype], _DTYPE[sampling_loc.dtype]))


def forward_index_bytes(value, sampling_loc) -> int:
    """msda_index_bytes for a call with these tensors."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_loc.shape
    return int(_lib.load().msda_index_bytes(N, S, M, D, L, Lq, P))


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step: int, flags: int | None = None, index=None, grads=None,
                            workspace=None) -> List[torch.Tensor]:
    _check_inputs([("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                   ("sampling_loc", sampling_loc), ("attn_weight", attn_weight), ("grad_output", grad_output)])
    dims, vdt, adt = _problem(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    N, S, M, D, L, Lq, P = dims
    _require(grad_output.dtype == value.dtype and grad_output.numel() == N * Lq * M * D,
             "grad_output must be (N, Lq, M*D) in the dtype of value")
    lib = _lib.load()
    flags = DEFAULT_FLAGS if flags is None else flags
    with torch.cuda.device(value.device):
        if grads is None:
            grad_value = torch.empty_like(value)
            grad_loc = torch.empty_like(sampling_loc)
            grad_attn = torch.empty_like(attn_weight)
        else:                                   # caller-owned result buffers
            grad_value = _into(grads[0], value.shape, value.dtype, value.device, "grads[0]")
            grad_loc = _into(grads[1], sampling_loc.shape, sampling_loc.dtype, value.device, "grads[1]")
            grad_attn = _into(grads[2], attn_weight.shape, attn_weight.dtype, value.device, "grads[2]")
        ws_bytes = 0 if flags & _lib.FLAG_ATOMIC_GRAD_VALUE else int(lib.msda_backward_workspace_bytes(*dims, vdt, adt))
        if workspace is None:
            ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=value.device)
        else:
            _require(workspace.is_cuda and workspace.dtype == torch.uint8 and workspace.is_contiguous()
                     and workspace.numel() >= ws_bytes, f"workspace must hold {ws_bytes} bytes")
            ws = workspace
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.msda_backward_indexed(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            sampling_loc.data_ptr(), attn_weight.data_ptr(), grad_output.data_ptr(),
            grad_value.data_ptr(), grad_loc.data_ptr(), grad_attn.data_ptr(),
            ws.data_ptr(), ws_bytes, None if index is None else index.data_ptr(),
            0 if index is None else index.numel(), *dims, vdt, adt, int(im2col_step), stream, flags))
        # `ws` may be released to the caching allocator here: reuse is ordered on this stream
    return [grad_value, grad_loc, grad_attn]


def _mask_ptr(padding_mask, value):
    """Device pointer of the (N, S) padding mask (bool / uint8, non-zero = padded pixel) or None."""
    if padding_mask is None:
        return None
    _require(isinstance(padding_mask, torch.Tensor) and padding_mask.is_cuda and padding_mask.device == value.device,
             "padding_mask must be a CUDA tensor on the device of value")
    _require(padding_mask.dtype in (torch.bool, torch.uint8) and tuple(padding_mask.shape) == tuple(value.shape[:2])
             and padding_mask.is_contiguous(), "padding_mask must be a contiguous bool (N, S) tensor")
    return padding_mask.data_ptr()


def ms_deform_attn_backward_fused(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                                  im2col_step: int, flags: int | None = None, index=None,
                                  padding_mask=None) -> List[torch.Tensor]:
    """Backward of ``ms_deform_attn_forward_fused`` (msda_backward_fused): takes the saved fp32 sampling locations /
    attention weights and returns ``[grad_value, grad_sampling_offsets, grad_attention_logits]`` -- the chain rule
    of the module's softmax and location arithmetic is applied inside the sample-gradient kernel."""
    _check_inputs([("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                   ("sampling_loc", sampling_loc), ("attn_weight", attn_weight), ("grad_output", grad_output)])
    dims, vdt, adt = _problem(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    N, S, M, D, L, Lq, P = dims
    _require(grad_output.dtype == value.dtype and grad_output.numel() == N * Lq * M * D,
             "grad_output must be (N, Lq, M*D) in the dtype of value")
    lib = _lib.load()
    with torch.cuda.device(value.device):
        grad_value = torch.empty_like(value)
        grad_off = torch.empty_like(sampling_loc)
        grad_logits = torch.empty_like(attn_weight)
        ws_bytes = int(lib.msda_backward_workspace_bytes(*dims, vdt, adt))
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=value.device)
        _lib.check(lib.msda_backward_fused(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), _mask_ptr(padding_mask, value),
            sampling_loc.data_ptr(), attn_weight.data_ptr(), grad_output.data_ptr(),
            grad_value.data_ptr(), grad_off.data_ptr(), grad_logits.data_ptr(),
            ws.data_ptr(), ws_bytes, None if index is None else index.data_ptr(),
            0 if index is None else index.numel(), *dims, vdt, adt, int(im2col_step),
            torch.cuda.current_stream().cuda_stream, DEFAULT_FLAGS if flags is None else flags))
    return [grad_value, grad_off, grad_logits]


def ms_deform_attn_backward_fused_raw(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                                      attn_logits, grad_output, im2col_step: int, index, flags: int | None = None,
                                      padding_mask=None) -> List[torch.Tensor]:
    """Backward of ``ms_deform_attn_forward_fused(..., materialize=False)`` (msda_backward_fused_raw): takes what that
    forward took plus its index, returns ``[grad_value, grad_sampling_offsets, grad_attention_logits]`` in the dtypes
    of value / the raw projections."""
    _check_inputs([("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                   ("reference_points", reference_points), ("sampling_offsets", sampling_offsets),
                   ("attn_logits", attn_logits), ("grad_output", grad_output)])
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_offsets.shape
    dims = (N, S, M, D, L, Lq, P)
    _require(tuple(reference_points.shape) == (N, Lq, L, 2) and reference_points.dtype == torch.float32,
             "reference_points must be fp32 (N, Lq, L, 2)")
    _require(sampling_offsets.dtype == attn_logits.dtype and sampling_offsets.dtype in (value.dtype, torch.float32),
             "sampling_offsets / attn_logits must share a dtype: the value dtype or float32")
    _require(grad_output.dtype == value.dtype and grad_output.numel() == N * Lq * M * D,
             "grad_output must be (N, Lq, M*D) in the dtype of value")
    _require(index is not None, "the index of the matching forward is required")
    lib = _lib.load()
    with torch.cuda.device(value.device):
        grad_value = torch.empty_like(value)
        grad_off = torch.empty_like(sampling_offsets)
        grad_logits = torch.empty_like(attn_logits)
        vdt, adt = _DTYPE[value.dtype], _DTYPE[sampling_offsets.dtype]
        ws_bytes = int(lib.msda_backward_workspace_bytes(*dims, vdt, adt))
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=value.device)
        _lib.check(lib.msda_backward_fused_raw(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), _mask_ptr(padding_mask, value),
            reference_points.data_ptr(),
            sampling_offsets.data_ptr(), attn_logits.data_ptr(), grad_output.data_ptr(), grad_value.data_ptr(),
            grad_off.data_ptr(), grad_logits.data_ptr(), ws.data_ptr(), ws_bytes, index.data_ptr(), index.numel(),
            *dims, vdt, adt, int(im2col_step), torch.cuda.current_stream().cuda_stream,
            DEFAULT_FLAGS if flags is None else flags))
    return [grad_value, grad_off, grad_logits]


def fused_prologue_supported(value, n_levels: int, n_points: int, ref_dim: int) -> bool:
    """Whether ms_deform_attn_forward_fused has a kernel for this call: tile-kernel
    shapes (fp32 rows of 64/128/256 B, bf16 rows of 64/128/256 B, P in {4, 8}), L*P <= 16, 2-d reference
    points.  Anything else keeps the unfused sequence."""
    if not value.is_cuda or ref_dim != 2 or n_levels * n_points > 16 or n_points not in (4, 8):
        return False
    row = value.shape[-1] * value.element_size()
    return value.dtype in (torch.float32, torch.bfloat16) and row in (64, 128, 256)


def ms_deform_attn_forward_fused(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                                 attn_logits, im2col_step: int, flags: int | None = None, want_index: bool = False,
                                 materialize: bool = True, padding_mask=None):
    """Forward with the module's prologue fused in (msda_forward_fused): returns
    ``(output, sampling_locations fp32, attention_weights fp32[, index])``.  ``materialize=False``: the two middle
    results are never written (returned as None); the matching backward is ``ms_deform_attn_backward_fused_raw``.
    ``padding_mask`` (bool (N, S), True on padded pixels): the module's ``value.masked_fill(mask[..., None], 0)``
    applied inside the kernels -- pass the UNMASKED value and give the same mask to the backward."""
    _check_inputs([("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                   ("reference_points", reference_points), ("sampling_offsets", sampling_offsets),
                   ("attn_logits", attn_logits)])
    _require(value.dim() == 4 and sampling_offsets.dim() == 6 and sampling_offsets.shape[-1] == 2, "bad shapes")
    N, S, M, D = value.shape
    _, Lq, M2, L, P, _ = sampling_offsets.shape
    _require(M2 == M and spatial_shapes.shape[0] == L and attn_logits.numel() == N * Lq * M * L * P,
             "sampling_offsets / attn_logits do not match value / spatial_shapes")
    _require(tuple(reference_points.shape) == (N, Lq, L, 2) and reference_points.dtype == torch.float32,
             "reference_points must be fp32 (N, Lq, L, 2)")
    _require(sampling_offsets.dtype == attn_logits.dtype and
             (sampling_offsets.dtype == value.dtype or sampling_offsets.dtype == torch.float32),
             "sampling_offsets / attn_logits must share a dtype: the value dtype or float32")
    dims = (N, S, M, D, L, Lq, P)
    lib = _lib.load()
    with torch.cuda.device(value.device):
        out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        loc = torch.empty((N, Lq, M, L, P, 2), dtype=torch.float32, device=value.device) if materialize else None
        attn = torch.empty((N, Lq, M, L, P), dtype=torch.float32, device=value.device) if materialize else None
        index, index_ptr, index_bytes = None, None, 0
        if want_index:
            index_bytes = int(lib.msda_index_bytes(*dims))
            if index_bytes:
                index = torch.empty(index_bytes, dtype=torch.uint8, device=value.device)
                index_ptr = index.data_ptr()
        _lib.check(lib.msda_forward_fused(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), _mask_ptr(padding_mask, value),
            reference_points.data_ptr(),
            sampling_offsets.data_ptr(), attn_logits.data_ptr(), out.data_ptr(),
            None if loc is None else loc.data_ptr(), None if attn is None else attn.data_ptr(),
            index_ptr, index_bytes, *dims, _DTYPE[value.dtype], _DTYPE[sampling_offsets.dtype], int(im2col_step),
            torch.cuda.current_stream().cuda_stream, DEFAULT_FLAGS if flags is None else flags))
    return (out, loc, attn, index) if want_index else (out, loc, attn)


def last_launch_count() -> int:
    return int(_lib.load().msda_last_launch_count())


# --------------------------------------------------------------------------------------------
# residual add + LayerNorm in one pass each way (msda_add_layernorm_*, include/msda_b200.h)
def add_layernorm_supported(x: torch.Tensor) -> bool:
    return x.is_cuda and x.shape[-1] == 256 and x.dtype in (torch.float32, torch.bfloat16)


def add_layernorm_forward(branch: torch.Tensor, residual: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                          eps: float = 1e-5):
    """``LayerNorm(branch + residual) * gamma + beta`` over the last dim (256 channels).  Returns
    ``(out, presum, mean, rstd)``; ``presum`` (= branch + residual, what the backward needs) is written over
    ``branch``, which the caller must no longer use."""
    _check_inputs([("branch", branch), ("residual", residual), ("gamma", gamma), ("beta", beta)])
    _require(branch.shape == residual.shape and branch.dtype == residual.dtype, "branch and residual must match")
    _require(gamma.dtype == torch.float32 and beta.dtype == torch.float32 and gamma.numel() == branch.shape[-1]
             and beta.numel() == branch.shape[-1], "gamma / beta must be fp32 vectors of the channel count")
    _require(branch.dtype in _DTYPE, f"unsupported dtype {branch.dtype}")
    C = branch.shape[-1]
    rows = branch.numel() // C
    lib = _lib.load()
    with torch.cuda.device(branch.device):
        out = torch.empty_like(branch)
        mean = torch.empty(rows, dtype=torch.float32, device=branch.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=branch.device)
        _lib.check(lib.msda_add_layernorm_forward(
            branch.data_ptr(), residual.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), branch.data_ptr(),
            mean.data_ptr(), rstd.data_ptr(), rows, C, _DTYPE[branch.dtype], float(eps), torch.cuda.current_stream().cuda_stream))
    return out, branch, mean, rstd


def add_layernorm_backward(grad_out: torch.Tensor, presum: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor,
                           gamma: torch.Tensor):
    """-> ``(grad_in, grad_gamma, grad_beta)``; ``grad_in`` is the gradient of both the branch and the residual."""
    _check_inputs([("grad_out", grad_out), ("presum", presum), ("mean", mean), ("rstd", rstd), ("gamma", gamma)])
    _require(grad_out.shape == presum.shape and grad_out.dtype == presum.dtype, "grad_out and presum must match")
    C = presum.shape[-1]
    rows = presum.numel() // C
    lib = _lib.load()
    with torch.cuda.device(presum.device):
        grad_in = torch.empty_like(presum)
        grad_gamma = torch.empty(C, dtype=torch.float32, device=presum.device)
        grad_beta = torch.empty(C, dtype=torch.float32, device=presum.device)
        ws_bytes = int(lib.msda_add_layernorm_backward_workspace_bytes(rows, C))
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=presum.device)
        _lib.check(lib.msda_add_layernorm_backward(
            grad_out.data_ptr(), presum.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), grad_in.data_ptr(),
            grad_gamma.data_ptr(), grad_beta.data_ptr(), ws.data_ptr(), ws_bytes, rows, C, _DTYPE[presum.dtype],
            torch.cuda.current_stream().cuda_stream))
    return grad_in, grad_gamma, grad_beta
