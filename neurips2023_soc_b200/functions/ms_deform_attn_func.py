"""``MSDeformAttnFunction`` -- the autograd boundary of the op, same name and call
signature as /root/reference/models/ops/functions/ms_deform_attn_func.py:21-38:

    MSDeformAttnFunction.apply(value, value_spatial_shapes, value_level_start_index,
                               sampling_locations, attention_weights, im2col_step)

Forward calls ``ms_deform_attn_forward`` and saves the five input tensors (not the
output, :27); backward is once-differentiable (:31) and returns gradients for value,
sampling_locations and attention_weights, ``None`` for the rest (:38).

Beyond the reference (fp32/fp64 only, ms_deform_attn_cuda.cu:64) bf16/fp16 values are
accepted; locations and weights may then stay fp32, the mix autocast produces.  If the
auxiliary tensors arrive in any other dtype they are promoted to fp32 here and their
gradients cast back.
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import msda_ext


def _aux_dtype(value, loc, attn):
    if loc.dtype == attn.dtype and (loc.dtype == value.dtype or
                                    (loc.dtype == torch.float32 and value.dtype != torch.float64)):
        return loc.dtype
    return torch.float64 if value.dtype == torch.float64 else torch.float32


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        aux = _aux_dtype(value, sampling_locations, attention_weights)
        ctx.aux_in = (sampling_locations.dtype, attention_weights.dtype)
        loc = sampling_locations.to(aux).contiguous()
        attn = attention_weights.to(aux).contiguous()
        value = value.contiguous()
        # when a backward will follow, the forward also leaves the sub-bin offsets of the
        # grad_value gather's inverse index (a pure function of the saved inputs)
        want_index = any(ctx.needs_input_grad[i] for i in (0, 3, 4))
        res = msda_ext.ms_deform_attn_forward(
            value, value_spatial_shapes, value_level_start_index, loc, attn, ctx.im2col_step, want_index=want_index)
        output, ctx.index = res if want_index else (res, None)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, loc, attn)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        grad_value, grad_loc, grad_attn = msda_ext.ms_deform_attn_backward(
            value, shapes, lsi, loc, attn, grad_output.to(value.dtype).contiguous(), ctx.im2col_step,
            index=ctx.index)
        ctx.index = None
        return grad_value, None, None, grad_loc.to(ctx.aux_in[0]), grad_attn.to(ctx.aux_in[1]), None
