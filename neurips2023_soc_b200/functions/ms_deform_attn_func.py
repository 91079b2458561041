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


class MSDeformAttnFusedFunction(Function):
    """The op together with the elementwise prologue of ``MSDeformAttn.forward``
    (/root/reference/models/ops/modules/ms_deform_attn.py:99-106, 2-d reference points):

        MSDeformAttnFusedFunction.apply(value, value_spatial_shapes, value_level_start_index,
                                        reference_points, sampling_offsets, attention_logits, im2col_step,
                                        padding_mask=None)
            -> (output, sampling_locations, attention_weights)

    ``padding_mask`` (bool (N, S), True on padded pixels): ``value`` is then the UNMASKED projection and the module's
    ``value.masked_fill(mask[..., None], 0)`` (:96-97) happens inside the kernels, forward and backward.

    The forward kernel forms ``softmax(attention_logits)`` and ``reference_points + sampling_offsets / (W, H)``
    in its staging threads (one pass over the raw projections instead of five elementwise kernels) and writes
    both out in fp32, as autocast would produce them.  The backward runs the same kernels as
    ``MSDeformAttnFunction`` on those saved tensors; the chain rule of the two elementwise maps is applied inside the
    sample-gradient kernel (msda_backward_fused) unless someone differentiated through the returned tensors.
    Only for shapes ``msda_ext.fused_prologue_supported`` accepts.
    """

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, reference_points, sampling_offsets,
                attention_logits, im2col_step, padding_mask=None):
        ctx.im2col_step = im2col_step
        ctx.padding_mask = padding_mask
        ctx.in_dtypes = (sampling_offsets.dtype, attention_logits.dtype, reference_points.dtype)
        # undefined gradients of the returned locations / weights must arrive as None, not as zero tensors:
        # that is how backward() knows nobody differentiated through them and takes the in-kernel chain rule
        ctx.set_materialize_grads(False)
        raw = sampling_offsets.dtype if (sampling_offsets.dtype == attention_logits.dtype and
                                         sampling_offsets.dtype in (value.dtype, torch.float32)) else torch.float32
        value = value.contiguous()
        want_index = any(ctx.needs_input_grad[i] for i in (0, 3, 4, 5))
        res = msda_ext.ms_deform_attn_forward_fused(
            value, value_spatial_shapes, value_level_start_index, reference_points.float().contiguous(),
            sampling_offsets.to(raw).contiguous(), attention_logits.to(raw).contiguous(), im2col_step,
            want_index=want_index, padding_mask=padding_mask)
        output, loc, attn = res[:3]
        ctx.index = res[3] if want_index else None
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, loc, attn)
        return output, loc, attn

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output, grad_loc_out, grad_attn_out):
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        if grad_output is None:               # only the returned locations / weights were differentiated
            grad_output = value.new_zeros((value.shape[0], loc.shape[1], value.shape[2] * value.shape[3]))
        go = grad_output.to(value.dtype).contiguous()
        index, ctx.index = ctx.index, None
        mask = ctx.padding_mask
        if grad_loc_out is None and grad_attn_out is None:
            # nobody differentiated through the returned locations / weights (the encoder drops them, the decoder
            # only ranks them): the chain rule of softmax and of ref + off / (W, H) runs inside the kernel
            grad_value, grad_offsets, grad_logits = msda_ext.ms_deform_attn_backward_fused(
                value, shapes, lsi, loc, attn, go, ctx.im2col_step, index=index, padding_mask=mask)
            grad_ref = None
            if ctx.needs_input_grad[3]:
                wh = shapes.flip(-1).to(grad_offsets.dtype)[None, None, None, :, None, :]
                grad_ref = (grad_offsets * wh).sum(dim=(2, 4)).to(ctx.in_dtypes[2])
            return (grad_value, None, None, grad_ref, grad_offsets.to(ctx.in_dtypes[0]),
                    grad_logits.flatten(-2).to(ctx.in_dtypes[1]), None, None)
        if mask is not None:                  # the unfused kernels know no mask: apply it around them
            value = value.masked_fill(mask[..., None, None], 0)
        grad_value, grad_loc, grad_attn = msda_ext.ms_deform_attn_backward(
            value, shapes, lsi, loc, attn, go, ctx.im2col_step, index=index)
        if mask is not None:
            grad_value = grad_value.masked_fill(mask[..., None, None], 0)
        if grad_loc_out is not None:          # someone differentiated through the returned locations / weights
            grad_loc = grad_loc + grad_loc_out
        if grad_attn_out is not None:
            grad_attn = grad_attn + grad_attn_out
        wh = shapes.flip(-1).to(grad_loc.dtype)[None, None, None, :, None, :]
        grad_offsets = (grad_loc / wh).to(ctx.in_dtypes[0])
        grad_ref = grad_loc.sum(dim=(2, 4)).to(ctx.in_dtypes[2]) if ctx.needs_input_grad[3] else None
        dot = (grad_attn * attn).sum(dim=(-1, -2), keepdim=True)
        grad_logits = (attn * (grad_attn - dot)).flatten(-2).to(ctx.in_dtypes[1])
        return grad_value, None, None, grad_ref, grad_offsets, grad_logits, None, None
