"""the fused encoder layer against the reference's op-by-op sequence.

* msda_add_layernorm_* against torch (add + F.layer_norm and autograd through them), fp32 tight, bf16 at bf16 accuracy;
* DeformableTransformerEncoderLayer with fused = True, compute_dtype fp32: outputs, input gradients and every
  parameter gradient against the same module with fused = False (the reference's sequence,
  /root/reference/models/deformable_transformer.py:253-263, on this repo's MSDeformAttn), with and without a
  padding mask;
* compute_dtype bf16 against the unfused layer under bf16 autocast;
* when the reference sources are staged (baseline/_ref/soc): a three-layer encoder of these layers, loaded from the
  state dict of the reference's own DeformableTransformerEncoder, against that encoder.
"""
import importlib
import sys
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

from neurips2023_soc_b200 import DeformableTransformerEncoder, DeformableTransformerEncoderLayer, msda_ext
from neurips2023_soc_b200.synthetic import level_start_index

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = Path(__file__).resolve().parent.parent
STAGED = ROOT / "baseline" / "_ref" / "soc"


def _rel(a, b):
    return float((a.double() - b.double()).abs().max()) / max(1.0, float(b.double().abs().max()))


@pytest.mark.parametrize("dt,tol", [(torch.float32, 2e-6), (torch.bfloat16, 2e-2)])
def test_add_layernorm_against_torch(dt, tol):
    g = torch.Generator().manual_seed(0)
    rows = 4099                                              # not a multiple of the rows a CTA holds
    a = torch.randn(rows, 256, generator=g).to(DEV, dt)
    b = (torch.randn(rows, 256, generator=g) * 2 + 0.5).to(DEV, dt)
    gamma = (torch.rand(256, generator=g) + 0.5).to(DEV)
    beta = torch.randn(256, generator=g).to(DEV)
    dy = torch.randn(rows, 256, generator=g).to(DEV, dt)
    a_ref = a.float().clone().requires_grad_(True)
    b_ref = b.float().clone().requires_grad_(True)
    g_ref, be_ref = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y_ref = F.layer_norm(a_ref + b_ref, (256,), g_ref, be_ref, 1e-5)
    y_ref.backward(dy.float())
    y, s, mean, rstd = msda_ext.add_layernorm_forward(a.clone(), b, gamma, beta, 1e-5)
    assert _rel(y, y_ref) <= tol
    assert _rel(s, (a.float() + b.float())) <= tol
    dx, dgamma, dbeta = msda_ext.add_layernorm_backward(dy, s, mean, rstd, gamma)
    assert _rel(dx, a_ref.grad) <= tol and _rel(dx, b_ref.grad) <= tol
    assert _rel(dgamma, g_ref.grad) <= (2e-5 if dt == torch.float32 else tol)
    assert _rel(dbeta, be_ref.grad) <= (2e-5 if dt == torch.float32 else tol)
    dx2, dgamma2, dbeta2 = msda_ext.add_layernorm_backward(dy, s, mean, rstd, gamma)
    assert torch.equal(dx, dx2) and torch.equal(dgamma, dgamma2) and torch.equal(dbeta, dbeta2)   # fixed summation order
    with pytest.raises(RuntimeError):
        msda_ext.add_layernorm_forward(a[:, :128].contiguous(), b[:, :128].contiguous(), gamma[:128], beta[:128])


def _layer_inputs(N, shapes_l, seed, pad):
    g = torch.Generator().manual_seed(seed)
    S = sum(h * w for h, w in shapes_l)
    src = torch.randn(N, S, 256, generator=g).to(DEV)
    pos = torch.randn(N, S, 256, generator=g).to(DEV)
    shapes = torch.tensor(shapes_l, dtype=torch.long, device=DEV)
    lsi = torch.tensor(level_start_index(shapes_l), dtype=torch.long, device=DEV)
    ratios = torch.ones(N, len(shapes_l), 2, device=DEV)
    ref = DeformableTransformerEncoder.get_reference_points(shapes, ratios, DEV)
    mask = None
    if pad:
        mask = torch.zeros(N, S, dtype=torch.bool, device=DEV)
        mask[1, -37:] = True
    up = torch.randn(N, S, 256, generator=g).to(DEV)
    return src, pos, ref, shapes, lsi, mask, up


def _make_layer(seed=0, d_ffn=512):
    torch.manual_seed(seed)
    layer = DeformableTransformerEncoderLayer(256, d_ffn, 0.0, "relu", 4, 8, 4).to(DEV)
    with torch.no_grad():                                    # leave the all-zero init of the offset / weight projections
        layer.self_attn.sampling_offsets.weight.normal_(0, 0.02)
        layer.self_attn.attention_weights.weight.normal_(0, 0.1)
        layer.norm1.weight.uniform_(0.5, 1.5)
        layer.norm2.bias.normal_(0, 0.1)
    return layer


def _run_layer(layer, fused, dtype, inputs, autocast=False):
    src, pos, ref, shapes, lsi, mask, up = inputs
    layer.fused, layer.compute_dtype = fused, dtype
    layer.zero_grad(set_to_none=True)
    s = src.clone().requires_grad_(True)
    p = pos.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        out = layer(s, p, ref, shapes, lsi, mask)
    (out.float() * up).sum().backward()
    grads = {n: q.grad.detach().clone() for n, q in layer.named_parameters()}
    return out.detach().float(), s.grad.detach().float(), p.grad.detach().float(), grads


@pytest.mark.parametrize("pad", [False, True])
def test_fused_layer_matches_reference_sequence_fp32(pad):
    layer = _make_layer()
    inputs = _layer_inputs(2, [(12, 20), (6, 10), (3, 5), (2, 3)], seed=1, pad=pad)
    a = _run_layer(layer, True, torch.float32, inputs)
    b = _run_layer(layer, False, torch.float32, inputs)
    assert _rel(a[0], b[0]) <= 2e-5 and _rel(a[1], b[1]) <= 5e-5 and _rel(a[2], b[2]) <= 5e-5
    assert a[3].keys() == b[3].keys() and len(a[3]) == 16
    for n in a[3]:
        assert _rel(a[3][n], b[3][n]) <= 1e-4, n


def test_fused_layer_bf16_against_autocast_reference_sequence():
    layer = _make_layer(seed=3)
    inputs = _layer_inputs(2, [(12, 20), (6, 10), (3, 5), (2, 3)], seed=4, pad=False)
    a = _run_layer(layer, True, torch.bfloat16, inputs)
    b = _run_layer(layer, False, torch.float32, inputs)      # fp32 truth
    c = _run_layer(layer, False, torch.float32, inputs, autocast=True)
    # the fused bf16 layer is as close to the fp32 layer as the autocast layer is (within a factor)
    for i in range(3):
        assert _rel(a[i], b[i]) <= max(5e-2, 3 * _rel(c[i], b[i])), i
    for n in a[3]:
        assert _rel(a[3][n], b[3][n]) <= max(5e-2, 3 * _rel(c[3][n], b[3][n])), n


def test_unsupported_settings_take_the_reference_sequence():
    layer = DeformableTransformerEncoderLayer(256, 128, 0.1, "gelu", 4, 8, 4).to(DEV)
    inputs = _layer_inputs(1, [(6, 10), (3, 5), (2, 3), (1, 2)], seed=5, pad=False)
    src, pos, ref, shapes, lsi, mask, up = inputs
    assert not layer._fusable(src, ref)                      # gelu
    layer2 = DeformableTransformerEncoderLayer(256, 128, 0.1, "relu", 4, 8, 4).to(DEV)
    layer2.train()
    assert not layer2._fusable(src, ref)                     # active dropout
    layer2.eval()
    assert layer2._fusable(src, ref)
    out = layer(src, pos, ref, shapes, lsi, mask)
    assert out.shape == src.shape and out.dtype == src.dtype


def test_encoder_of_fused_layers_against_reference_encoder():
    if not (STAGED / "MANIFEST.json").exists():
        pytest.skip("reference sources not staged")
    saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.") or k == "misc"}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, str(STAGED))
    sys.path.insert(0, str(ROOT))
    try:
        dt = importlib.import_module("models.deformable_transformer")
        torch.manual_seed(0)
        ref_enc = dt.DeformableTransformerEncoder(dt.DeformableTransformerEncoderLayer(256, 512, 0.0, "relu", 4, 8, 4), 3).to(DEV)
        with torch.no_grad():
            for n, p in ref_enc.named_parameters():
                if "sampling_offsets.weight" in n:
                    p.normal_(0, 0.02)
                if "attention_weights.weight" in n:
                    p.normal_(0, 0.1)
        ours = DeformableTransformerEncoder(DeformableTransformerEncoderLayer(256, 512, 0.0, "relu", 4, 8, 4), 3).to(DEV)
        ours.load_state_dict(ref_enc.state_dict())           # same keys: the reference's checkpoint loads
        for layer in ours.layers:
            layer.compute_dtype = torch.float32
        shapes_l = [(12, 20), (6, 10), (3, 5), (2, 3)]
        src, pos, _, shapes, lsi, mask, up = _layer_inputs(2, shapes_l, seed=7, pad=True)
        ratios = torch.rand(2, 4, 2, device=DEV) * 0.2 + 0.8
        outs = []
        for enc in (ours, ref_enc):
            enc.zero_grad(set_to_none=True)
            out = enc(src, shapes, lsi, ratios, pos, mask)
            (out * up).sum().backward()
            outs.append((out.detach(), {n: p.grad.detach().clone() for n, p in enc.named_parameters()}))
        assert _rel(outs[0][0], outs[1][0]) <= 2e-5
        for n in outs[1][1]:
            assert _rel(outs[0][1][n], outs[1][1][n]) <= 2e-4, n
    finally:
        sys.path.remove(str(ROOT))
        sys.path.remove(str(STAGED))
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "misc"]:
            del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.parametrize("vdt,adt", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32)])
def test_raw_fused_backward_matches_materialised_one(vdt, adt):
    """msda_forward_fused without its two middle outputs + msda_backward_fused_raw against the materialising pair:
    same output bits, same grad_value bits (the index entries carry the same values), gradients of the raw
    projections equal up to the rounding of the dtype they are written in."""
    g = torch.Generator().manual_seed(11)
    shapes_l = [(12, 20), (6, 10), (3, 5), (2, 3)]
    S = sum(h * w for h, w in shapes_l)
    N, M, D, L, P = 2, 8, 32, 4, 4
    shapes = torch.tensor(shapes_l, dtype=torch.long, device=DEV)
    lsi = torch.tensor(level_start_index(shapes_l), dtype=torch.long, device=DEV)
    value = torch.randn(N, S, M, D, generator=g).to(DEV, vdt)
    ref = DeformableTransformerEncoder.get_reference_points(shapes_l, torch.ones(N, L, 2, device=DEV), DEV).contiguous()
    off = (torch.randn(N, S, M, L, P, 2, generator=g) * 3).to(DEV, adt)
    logit = torch.randn(N, S, M, L * P, generator=g).to(DEV, adt)
    go = torch.randn(N, S, M * D, generator=g).to(DEV, vdt)
    out_a, loc, attn, idx_a = msda_ext.ms_deform_attn_forward_fused(value, shapes, lsi, ref, off, logit, 64, want_index=True)
    out_b, none1, none2, idx_b = msda_ext.ms_deform_attn_forward_fused(value, shapes, lsi, ref, off, logit, 64, want_index=True,
                                                                       materialize=False)
    assert none1 is None and none2 is None and torch.equal(out_a, out_b) and torch.equal(idx_a, idx_b)
    gv_a, goff_a, glog_a = msda_ext.ms_deform_attn_backward_fused(value, shapes, lsi, loc, attn, go, 64, index=idx_a)
    gv_b, goff_b, glog_b = msda_ext.ms_deform_attn_backward_fused_raw(value, shapes, lsi, ref, off, logit, go, 64, index=idx_b)
    assert goff_b.dtype == adt and glog_b.dtype == adt
    assert torch.equal(gv_a, gv_b)
    tol = 1e-6 if adt == torch.float32 else 1e-2
    assert _rel(goff_b, goff_a) <= tol and _rel(glog_b.view_as(glog_a), glog_a) <= tol
    with pytest.raises(RuntimeError):          # a decoder-shaped call keeps no index: not this entry point
        msda_ext.ms_deform_attn_backward_fused_raw(value, shapes, lsi, ref[:, :5].contiguous(), off[:, :5].contiguous(),
                                                   logit[:, :5].contiguous(), go[:, :5].contiguous(), 64, index=idx_b)


@pytest.mark.parametrize("vdt,adt", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32), (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("Lq", [None, 20])
def test_padding_mask_inside_the_kernels_equals_masked_fill(vdt, adt, Lq):
    """The module's value.masked_fill(padding_mask, 0) (ms_deform_attn.py:96-97) applied inside the kernels -- padded
    pixels dropped corner by corner in the forward and the sample-gradient kernel, their grad_value rows zeroed by the
    walker / the direct gather -- against masked_fill around the same kernels: the same bits, for encoder-shaped calls
    (inverse index; materialising and raw backward) and decoder-shaped ones (Lq = 20: direct gather)."""
    g = torch.Generator().manual_seed(13)
    shapes_l = [(12, 20), (6, 10), (3, 5), (2, 3)]
    S = sum(h * w for h, w in shapes_l)
    N, M, D, L, P = 2, 8, 32, 4, 4
    shapes = torch.tensor(shapes_l, dtype=torch.long, device=DEV)
    lsi = torch.tensor(level_start_index(shapes_l), dtype=torch.long, device=DEV)
    value = torch.randn(N, S, M, D, generator=g).to(DEV, vdt)
    mask = (torch.rand(N, S, generator=g) < 0.3).to(DEV)                 # a third of the pixels are padding
    mask[0, :40] = True                                                  # whole rows of the finest level too
    nq = S if Lq is None else Lq
    if Lq is None:
        ref = DeformableTransformerEncoder.get_reference_points(shapes_l, torch.ones(N, L, 2, device=DEV), DEV).contiguous()
    else:
        ref = torch.rand(N, nq, L, 2, generator=g).to(DEV)
    off = (torch.randn(N, nq, M, L, P, 2, generator=g) * 3).to(DEV, adt)
    logit = torch.randn(N, nq, M, L * P, generator=g).to(DEV, adt)
    go = torch.randn(N, nq, M * D, generator=g).to(DEV, vdt)
    filled = value.masked_fill(mask[..., None, None], 0)

    out_m, loc_m, attn_m, idx_m = msda_ext.ms_deform_attn_forward_fused(value, shapes, lsi, ref, off, logit, 64, want_index=True,
                                                                        padding_mask=mask)
    out_f, loc_f, attn_f, idx_f = msda_ext.ms_deform_attn_forward_fused(filled, shapes, lsi, ref, off, logit, 64, want_index=True)
    assert torch.equal(out_m, out_f) and torch.equal(loc_m, loc_f) and torch.equal(attn_m, attn_f)
    gm = msda_ext.ms_deform_attn_backward_fused(value, shapes, lsi, loc_m, attn_m, go, 64, index=idx_m, padding_mask=mask)
    gf = msda_ext.ms_deform_attn_backward_fused(filled, shapes, lsi, loc_f, attn_f, go, 64, index=idx_f)
    gf[0] = gf[0].masked_fill(mask[..., None, None], 0)                  # masked_fill's own backward
    for a, b in zip(gm, gf):
        assert torch.equal(a, b)
    assert float(gm[0][mask].abs().max()) == 0.0
    if Lq is None:                                                       # the raw pair (what the encoder layer runs)
        out_r, _, _, idx_r = msda_ext.ms_deform_attn_forward_fused(value, shapes, lsi, ref, off, logit, 64, want_index=True,
                                                                   materialize=False, padding_mask=mask)
        assert torch.equal(out_r, out_f)
        gr = msda_ext.ms_deform_attn_backward_fused_raw(value, shapes, lsi, ref, off, logit, go, 64, index=idx_r, padding_mask=mask)
        gfr = msda_ext.ms_deform_attn_backward_fused_raw(filled, shapes, lsi, ref, off, logit, go, 64,
                                                         index=msda_ext.ms_deform_attn_forward_fused(
                                                             filled, shapes, lsi, ref, off, logit, 64, want_index=True,
                                                             materialize=False)[3])
        gfr[0] = gfr[0].masked_fill(mask[..., None, None], 0)
        for a, b in zip(gr, gfr):
            assert torch.equal(a, b)
    with pytest.raises(RuntimeError, match="padding_mask"):
        msda_ext.ms_deform_attn_forward_fused(value, shapes, lsi, ref, off, logit, 64, padding_mask=mask[:, :-1])
