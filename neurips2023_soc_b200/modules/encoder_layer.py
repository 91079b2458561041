"""``DeformableTransformerEncoderLayer`` / ``DeformableTransformerEncoder`` -- the caller of the op in the encoder
(/root/reference/models/deformable_transformer.py:225-263 and :266-293), rebuilt around this repo's kernels
(SURVEY.md 8f-2).  Same constructor arguments, same parameter names (``self_attn.{sampling_offsets,
attention_weights, value_proj, output_proj}``, ``norm1``, ``linear1``, ``linear2``, ``norm2``), same ``forward``
signatures, so reference checkpoints load and the reference's ``DeformableTransformer`` can hold these layers.

What differs is how a layer runs on the B200 (``fused = True``, the default on CUDA for d_model = 256, ReLU, no
active dropout):

* the layer computes in ONE dtype end to end (``compute_dtype``: bf16 by default, fp32 for parity checks) -- under
  autocast the stock layer keeps the residual stream in fp32 and pays a cast around every GEMM;
* ``softmax`` and ``reference_points + offsets / (W, H)`` run inside the forward kernel, their chain rule inside the
  sample-gradient kernel (``msda_forward_fused`` / ``msda_backward_fused``);
* ``norm(src + branch)`` is one kernel forward and one backward (``msda_add_layernorm_*``) instead of add +
  LayerNorm + three backward kernels;
* the backward of the whole layer is written out by hand: bias gradients are GEMVs on the tensor cores instead of
  a strided reduction per Linear, residual gradients are folded into the GEMMs' accumulate (``addmm``), the ReLU
  into the first FFN GEMM's epilogue.
The dense contractions themselves stay with cuBLASLt through torch (integration-level win, not a
kernel of this repo).  Anything the fused path does not cover (other activations, active dropout, other widths,
CPU tensors) takes the reference's op-by-op sequence with this repo's ``MSDeformAttn`` inside.
"""
from __future__ import annotations

import copy

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import msda_ext
from .ms_deform_attn import MSDeformAttn


def _activation(name: str):
    if name == "relu":
        return F.relu
    if name == "gelu":
        return F.gelu
    if name == "glu":
        return F.glu
    raise RuntimeError(f"activation should be relu/gelu, not {name}.")


def _colsum(t2d: torch.Tensor) -> torch.Tensor:
    """Sum over rows as a GEMV (ones @ t2d) on the tensor cores: what a Linear's bias gradient is."""
    ones = torch.ones((1, t2d.shape[0]), dtype=t2d.dtype, device=t2d.device)
    return torch.mm(ones, t2d)[0]


class _EncoderLayerFunction(Function):
    """One encoder layer, forward and hand-written backward, in one compute dtype."""

    @staticmethod
    def forward(ctx, x, pos, ref, shapes, lsi, pad, n_heads, n_levels, n_points, eps1, eps2,
                Wso, bso, Waw, baw, Wv, bv, Wo, bo, g1, be1, W1, b1, W2, b2, g2, be2):
        dt = x.dtype
        N, S, C = x.shape
        M, L, P = n_heads, n_levels, n_points
        T = N * S
        cast = lambda w: w.to(dt)                                              # noqa: E731  (fp32 masters -> compute dtype)
        Wso_, Waw_, Wv_, Wo_, W1_, W2_ = (cast(w) for w in (Wso, Waw, Wv, Wo, W1, W2))
        x2 = x.reshape(T, C)
        q2 = x2 if pos is None else (x + pos).reshape(T, C)
        off = F.linear(q2, Wso_, cast(bso)).view(N, S, M, L, P, 2)
        logit = F.linear(q2, Waw_, cast(baw)).view(N, S, M, L * P)
        # the padding mask goes into the kernels (padded pixels count as zero rows; their gradient rows come back zero)
        value = F.linear(x2, Wv_, cast(bv)).view(N, S, M, C // M)
        if pad is not None:
            pad = pad.to(torch.bool).contiguous()
        # the encoder never looks at the sampling locations / attention weights: they are not written at all when the
        # call keeps an inverse index (many queries per frame), the backward then recomputes them from the raw projections
        raw = msda_ext.forward_index_bytes(value, off) > 0
        attn_out, loc, attn, index = msda_ext.ms_deform_attn_forward_fused(value, shapes, lsi, ref, off, logit, 64,
                                                                           want_index=True, materialize=not raw,
                                                                           padding_mask=pad)
        if raw:
            loc, attn = off, logit
        a = F.linear(attn_out.view(T, C), Wo_, cast(bo))
        x1, s1, mean1, rstd1 = msda_ext.add_layernorm_forward(a, x2, g1, be1, eps1)       # s1 overwrites a
        h = torch._addmm_activation(cast(b1), x1, W1_.t())                                 # relu(x1 W1^T + b1), one GEMM
        y = F.linear(h, W2_, cast(b2))
        out, s2, mean2, rstd2 = msda_ext.add_layernorm_forward(y, x1, g2, be2, eps2)       # s2 overwrites y
        ctx.save_for_backward(x2, q2, value, loc, attn, attn_out, s1, mean1, rstd1, x1, h, s2, mean2, rstd2,
                              shapes, lsi, Wso_, Waw_, Wv_, Wo_, W1_, W2_, g1, g2)
        ctx.index, ctx.pad, ctx.raw, ctx.ref = index, pad, raw, (ref if raw else None)
        ctx.has_pos = pos is not None
        ctx.dims = (N, S, C, M, L, P)
        return out.view(N, S, C)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        (x2, q2, value, loc, attn, attn_out, s1, mean1, rstd1, x1, h, s2, mean2, rstd2, shapes, lsi,
         Wso_, Waw_, Wv_, Wo_, W1_, W2_, g1, g2) = ctx.saved_tensors
        N, S, C, M, L, P = ctx.dims
        T = N * S
        dt = x2.dtype
        f32 = torch.float32
        dout = dout.to(dt).contiguous().view(T, C)
        # norm2(x1 + ffn(x1))
        dz2, dg2, dbe2 = msda_ext.add_layernorm_backward(dout, s2, mean2, rstd2, g2)
        dW2 = torch.mm(dz2.t(), h)
        db2 = _colsum(dz2)
        dh = torch.ops.aten.threshold_backward(torch.mm(dz2, W2_), h, 0.0)             # ReLU: pass where h > 0
        dW1 = torch.mm(dh.t(), x1)
        db1 = _colsum(dh)
        dx1 = torch.addmm(dz2, dh, W1_)                                         # residual + FFN branch
        del dh
        # norm1(x + output_proj(attn))
        dz1, dg1, dbe1 = msda_ext.add_layernorm_backward(dx1, s1, mean1, rstd1, g1)
        ao2 = attn_out.view(T, C)
        dWo = torch.mm(dz1.t(), ao2)
        dbo = _colsum(dz1)
        dao = torch.mm(dz1, Wo_).view(N, S, C)
        index, ctx.index = ctx.index, None
        if ctx.raw:      # loc / attn hold the raw offsets / logits
            dvalue, doff, dlogit = msda_ext.ms_deform_attn_backward_fused_raw(value, shapes, lsi, ctx.ref, loc, attn, dao, 64,
                                                                               index=index, padding_mask=ctx.pad)
        else:
            dvalue, doff, dlogit = msda_ext.ms_deform_attn_backward_fused(value, shapes, lsi, loc, attn, dao, 64, index=index,
                                                                          padding_mask=ctx.pad)
        dv2 = dvalue.reshape(T, C)
        dWv = torch.mm(dv2.t(), x2)
        dbv = _colsum(dv2)
        dx = torch.addmm(dz1, dv2, Wv_)                                         # residual + value branch
        doff2 = doff.reshape(T, M * L * P * 2).to(dt)
        dlog2 = dlogit.reshape(T, M * L * P).to(dt)
        dWso = torch.mm(doff2.t(), q2)
        dbso = _colsum(doff2)
        dWaw = torch.mm(dlog2.t(), q2)
        dbaw = _colsum(dlog2)
        dq = torch.addmm(torch.mm(doff2, Wso_), dlog2, Waw_)
        dx += dq                                                                # query branch (q = x + pos)
        dpos = dq.view(N, S, C) if (ctx.has_pos and ctx.needs_input_grad[1]) else None
        g = lambda t: t.to(f32)                                                 # noqa: E731  (gradients of the fp32 masters)
        return (dx.view(N, S, C), dpos, None, None, None, None, None, None, None, None, None,
                g(dWso), g(dbso), g(dWaw), g(dbaw), g(dWv), g(dbv), g(dWo), g(dbo), dg1, dbe1,
                g(dW1), g(db1), g(dW2), g(db2), dg2, dbe2)


class DeformableTransformerEncoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        # self attention
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        # ffn
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _activation(activation)
        self._activation_name = activation
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        #: run the layer through _EncoderLayerFunction when it applies (see the module docstring)
        self.fused = True
        #: dtype the fused layer computes in; the result is returned in it
        self.compute_dtype = torch.bfloat16

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def forward_ffn(self, src):
        src2 = self.linear2(self.dropout2(self.activation(self.linear1(src))))
        src = src + self.dropout3(src2)
        return self.norm2(src)

    def _fusable(self, src, reference_points) -> bool:
        a = self.self_attn
        if not (self.fused and src.is_cuda and self._activation_name == "relu"):
            return False
        if self.training and any(d.p > 0 for d in (self.dropout1, self.dropout2, self.dropout3)):
            return False
        if a.d_model != 256 or self.linear1.out_features % 8 or reference_points.shape[-1] != 2:
            return False
        probe = torch.empty((1, 1, a.n_heads, a.d_model // a.n_heads), dtype=self.compute_dtype, device=src.device)
        return msda_ext.fused_prologue_supported(probe, a.n_levels, a.n_points, 2)

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None):
        if not self._fusable(src, reference_points):
            # the reference's sequence (deformable_transformer.py:253-263)
            src2, _, _ = self.self_attn(self.with_pos_embed(src, pos), reference_points, src, spatial_shapes,
                                        level_start_index, padding_mask)
            src = self.norm1(src + self.dropout1(src2))
            return self.forward_ffn(src)
        a, dt = self.self_attn, self.compute_dtype
        a._check_geometry(spatial_shapes, src.shape[1])
        return _EncoderLayerFunction.apply(
            src.to(dt).contiguous(), None if pos is None else pos.to(dt).contiguous(),
            reference_points.float().contiguous(), spatial_shapes, level_start_index, padding_mask,
            a.n_heads, a.n_levels, a.n_points, self.norm1.eps, self.norm2.eps,
            a.sampling_offsets.weight, a.sampling_offsets.bias, a.attention_weights.weight, a.attention_weights.bias,
            a.value_proj.weight, a.value_proj.bias, a.output_proj.weight, a.output_proj.bias,
            self.norm1.weight, self.norm1.bias, self.linear1.weight, self.linear1.bias,
            self.linear2.weight, self.linear2.bias, self.norm2.weight, self.norm2.bias)


class DeformableTransformerEncoder(nn.Module):
    """deformable_transformer.py:266-293: reference points = the pixel centres of every level (scaled by the valid
    ratios), then the layers.  With fused layers the positions are cast once and the stream stays in the layers'
    compute dtype until the end, where it returns to the dtype of ``src``."""

    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device):
        """The reference's arithmetic (deformable_transformer.py:272-285); ``spatial_shapes`` may be a tensor (read on
        the host, as the reference does) or a list of (H, W) pairs."""
        pts = []
        for lvl, (H_, W_) in enumerate(spatial_shapes):
            H_, W_ = int(H_), int(W_)
            ref_y, ref_x = torch.meshgrid(torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32, device=device),
                                          torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32, device=device), indexing="ij")
            ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H_)
            ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W_)
            pts.append(torch.stack((ref_x, ref_y), -1))
        reference_points = torch.cat(pts, 1)
        return reference_points[:, :, None] * valid_ratios[:, None]

    def _host_shapes(self, spatial_shapes):
        """The level shapes as Python ints, read from the device once per distinct shapes tensor: the reference reads
        them on every call (`for lvl, (H_, W_) in enumerate(spatial_shapes)`), four host synchronisations per
        forward that keep the host from running ahead of the GPU."""
        key = (spatial_shapes.data_ptr(), spatial_shapes._version, tuple(spatial_shapes.shape))
        if getattr(self, "_shapes_key", None) != key:
            if spatial_shapes.is_cuda and torch.cuda.is_current_stream_capturing():
                raise RuntimeError("call the encoder once outside CUDA graph capture so that the level shapes are known")
            self._shapes_host = [tuple(map(int, hw)) for hw in spatial_shapes.tolist()]
            self._shapes_key = key
        return self._shapes_host

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None):
        out_dtype = src.dtype
        reference_points = self.get_reference_points(self._host_shapes(spatial_shapes), valid_ratios, device=src.device)
        output = src
        first = self.layers[0] if len(self.layers) else None
        if first is not None and getattr(first, "fused", False) and first._fusable(src, reference_points):
            output = src.to(first.compute_dtype)
            pos = None if pos is None else pos.to(first.compute_dtype)
        for layer in self.layers:
            output = layer(output, pos, reference_points, spatial_shapes, level_start_index, padding_mask)
        return output.to(out_dtype)
