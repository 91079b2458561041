"""Parity of the sm_100a kernels (through the C ABI) with the CPU oracle and the golden vectors.

Tolerances (BASELINE.json north_star): max-abs error <= 1e-5 for fp32 and <= 2e-2 for bf16,
taken relative to max(1, max|reference|) so that gradients that sum thousands of terms are
judged on the same footing as O(1) outputs.  bf16 cases compare against the oracle evaluated in
fp64 on the bf16-rounded inputs.
"""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from neurips2023_soc_b200 import MSDeformAttnFunction, _lib, msda_ext
from neurips2023_soc_b200.synthetic import A2D_PYRAMID, make_inputs
from oracle import msda_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
TOL = {torch.float64: 1e-11, torch.float32: 1e-5, torch.bfloat16: 2e-2, torch.float16: 2e-3}


def rel_err(a, ref, keep=None):
    a = a.detach().double().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    err = np.abs(a.reshape(ref.shape) - ref)
    if keep is not None:
        err = err * keep
    return float(err.max()) / max(1.0, float(np.abs(ref).max()))


def off_lattice(loc, shapes, eps=1e-3):
    loc = np.asarray(loc, np.float64)
    wh = np.asarray(shapes)[:, ::-1].astype(np.float64).reshape(1, 1, 1, -1, 1, 2)
    px = loc * wh - 0.5
    ok = (np.abs(px - np.round(px)) > eps).all(-1, keepdims=True)
    return np.broadcast_to(ok, loc.shape).astype(np.float64)


def run_op(value, shapes, lsi, loc, attn, grad_out, vdt, adt, flags=0):
    """forward + backward through the extension-level API on the GPU."""
    v = value.to(DEV, vdt).contiguous()
    lo, at = loc.to(DEV, adt).contiguous(), attn.to(DEV, adt).contiguous()
    sh, ls = shapes.to(DEV), lsi.to(DEV)
    go = grad_out.to(DEV, vdt).contiguous()
    out = msda_ext.ms_deform_attn_forward(v, sh, ls, lo, at, 64, flags=flags)
    gv, gl, ga = msda_ext.ms_deform_attn_backward(v, sh, ls, lo, at, go, 64, flags=flags)
    if not flags & _lib.FLAG_ATOMIC_GRAD_VALUE:
        # the index handed over by the forward must reproduce the self-counted backward bit for bit
        out2, index = msda_ext.ms_deform_attn_forward(v, sh, ls, lo, at, 64, flags=flags, want_index=True)
        gv2, gl2, ga2 = msda_ext.ms_deform_attn_backward(v, sh, ls, lo, at, go, 64, flags=flags, index=index)
        assert torch.equal(out, out2) and torch.equal(gl, gl2) and torch.equal(ga, ga2)
        assert (flags & _lib.FLAG_UNORDERED) or torch.equal(gv, gv2)     # unordered: same terms, arrival order
    torch.cuda.synchronize()
    return out, gv, gl, ga, (v, lo, at, go)


def oracle_f64(v, shapes, lsi, lo, at, go):
    args = [t.detach().double().cpu().numpy() for t in (v, lo, at, go)]
    out = O.forward_c(args[0], shapes.numpy(), lsi.numpy(), args[1], args[2])
    gv, gl, ga = O.backward_c(args[0], shapes.numpy(), lsi.numpy(), args[1], args[2], args[3])
    return out, gv, gl, ga


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("vdt,adt", [(torch.float32, torch.float32), (torch.float64, torch.float64)])
def test_golden_vectors(name, vdt, adt):
    g = load_golden(name)
    t = {k: torch.from_numpy(g[k]) for k in ("value", "loc", "attn", "grad_out", "shapes", "lsi")}
    out, gv, gl, ga, _ = run_op(t["value"], t["shapes"], t["lsi"], t["loc"], t["attn"], t["grad_out"], vdt, adt)
    tol = TOL[vdt]
    keep = off_lattice(g["loc"], g["shapes"], 1e-3 if vdt == torch.float32 else 0.0)
    if vdt == torch.float64:  # only the accept-boundary kink differs from autograd (see oracle header)
        wh = g["shapes"][:, ::-1].astype(np.float64).reshape(1, 1, 1, -1, 1, 2)
        px = g["loc"].astype(np.float64) * wh - 0.5
        keep = np.broadcast_to((px != -1.0).all(-1, keepdims=True), px.shape).astype(np.float64)
    assert rel_err(out, g["out"]) <= tol
    assert rel_err(gv, g["grad_value"]) <= tol
    assert rel_err(ga, g["grad_attn"]) <= tol
    assert rel_err(gl, g["grad_loc"], keep) <= 2 * tol


@pytest.mark.parametrize("name", ["encoder_like", "decoder_like", "p8_l1", "channels_64", "out_of_range"])
@pytest.mark.parametrize("vdt,adt", [(torch.bfloat16, torch.float32), (torch.bfloat16, torch.bfloat16),
                                      (torch.float16, torch.float32)])
def test_golden_inputs_reduced_precision(name, vdt, adt):
    g = load_golden(name)
    t = {k: torch.from_numpy(g[k]) for k in ("value", "loc", "attn", "grad_out", "shapes", "lsi")}
    out, gv, gl, ga, (v, lo, at, go) = run_op(t["value"], t["shapes"], t["lsi"], t["loc"], t["attn"],
                                               t["grad_out"], vdt, adt)
    r_out, r_gv, r_gl, r_ga = oracle_f64(v, t["shapes"], t["lsi"], lo, at, go)
    tol = TOL[vdt]
    keep = off_lattice(lo.double().cpu().numpy(), g["shapes"], 2e-2 if adt != torch.float32 else 1e-3)
    assert rel_err(out, r_out) <= tol
    assert rel_err(gv, r_gv) <= tol
    assert rel_err(ga, r_ga) <= tol
    assert rel_err(gl, r_gl, keep) <= 2 * tol


@pytest.mark.parametrize("dist", ["encoder", "uniform"])
@pytest.mark.parametrize("vdt,adt", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32)])
def test_config1_shape_vs_c_oracle(dist, vdt, adt):
    """BASELINE config 1: one frame of the A2D pyramid, 5100 queries, 8 heads x 32, 4 levels x 4 points."""
    x = make_inputs(N=1, dist=dist, seed=1)
    out, gv, gl, ga, (v, lo, at, go) = run_op(x.value, x.spatial_shapes, x.level_start_index,
                                               x.sampling_locations, x.attention_weights, x.grad_output, vdt, adt)
    r_out, r_gv, r_gl, r_ga = oracle_f64(v, x.spatial_shapes, x.level_start_index, lo, at, go)
    tol = TOL[vdt]
    keep = off_lattice(lo.double().cpu().numpy(), x.spatial_shapes.numpy(), 1e-3)
    assert rel_err(out, r_out) <= tol
    assert rel_err(gv, r_gv) <= tol
    assert rel_err(ga, r_ga) <= tol
    assert rel_err(gl, r_gl, keep) <= 2 * tol


@pytest.mark.parametrize("kw", [dict(N=2, dist="decoder", Lq=20), dict(N=3, dist="decoder", Lq=5),
                                dict(N=2, dist="uniform", Lq=300, P=8),
                                dict(N=1, dist="encoder", shapes=[(9, 13), (5, 7)], M=4, D=16),
                                dict(N=1, dist="encoder", shapes=[(9, 13), (5, 7)], M=2, D=64),
                                dict(N=1, dist="uniform", shapes=[(7, 5)], M=3, D=20, Lq=33, P=3)])
def test_shape_sweep_fp32(kw):
    x = make_inputs(seed=5, **kw)
    out, gv, gl, ga, (v, lo, at, go) = run_op(x.value, x.spatial_shapes, x.level_start_index,
                                               x.sampling_locations, x.attention_weights, x.grad_output,
                                               torch.float32, torch.float32)
    r_out, r_gv, r_gl, r_ga = oracle_f64(v, x.spatial_shapes, x.level_start_index, lo, at, go)
    keep = off_lattice(lo.double().cpu().numpy(), x.spatial_shapes.numpy(), 1e-3)
    assert rel_err(out, r_out) <= 1e-5
    assert rel_err(gv, r_gv) <= 1e-5
    assert rel_err(ga, r_ga) <= 1e-5
    assert rel_err(gl, r_gl, keep) <= 2e-5


def test_generic_and_tile_paths_agree():
    x = make_inputs(N=2, dist="encoder", shapes=[(12, 20), (6, 10), (3, 5), (2, 3)], seed=2)
    a = run_op(x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights,
               x.grad_output, torch.float32, torch.float32)
    b = run_op(x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights,
               x.grad_output, torch.float32, torch.float32, flags=_lib.FLAG_GENERIC)
    c = run_op(x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights,
               x.grad_output, torch.float32, torch.float32, flags=_lib.FLAG_PYRAMID_TILES)
    for i in range(4):
        assert rel_err(a[i], b[i].double().cpu().numpy()) <= 1e-5
        assert torch.equal(a[i], c[i]), "query tiling must not change a single bit"


def test_backward_is_bit_reproducible():
    """The reference's fp32 atomicAdd scatter (cuh:125-152) is order-dependent; this one is not."""
    x = make_inputs(N=2, dist="encoder", seed=3)
    first = None
    for _ in range(5):
        r = run_op(x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights,
                   x.grad_output, torch.float32, torch.float32)
        if first is None:
            first = r
        else:
            for i in range(4):
                assert torch.equal(first[i], r[i])


def test_atomic_arm_matches_deterministic_path():
    x = make_inputs(N=1, dist="encoder", seed=4)
    a = run_op(x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights,
               x.grad_output, torch.float32, torch.float32)
    b = run_op(x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights,
               x.grad_output, torch.float32, torch.float32, flags=_lib.FLAG_ATOMIC_GRAD_VALUE)
    for i in range(4):
        assert rel_err(b[i], a[i].double().cpu().numpy()) <= 1e-5


def test_many_samples_in_one_bin():
    """All queries look at the same spot: one bin holds every sample of its level (exercises the
    shared-memory and the in-place global sort of big bins)."""
    x = make_inputs(N=1, dist="uniform", shapes=[(6, 10), (3, 5)], M=2, D=32, Lq=1500, seed=6)
    loc = x.sampling_locations * 0.0 + torch.tensor([0.43, 0.57])
    loc[:, :700] += 0.001 * torch.randn(1, 700, 2, 2, 4, 2, generator=torch.Generator().manual_seed(1))
    res = [run_op(x.value, x.spatial_shapes, x.level_start_index, loc, x.attention_weights, x.grad_output,
                  torch.float32, torch.float32) for _ in range(2)]
    out, gv, gl, ga, (v, lo, at, go) = res[0]
    r_out, r_gv, r_gl, r_ga = oracle_f64(v, x.spatial_shapes, x.level_start_index, lo, at, go)
    assert rel_err(out, r_out) <= 1e-5
    assert rel_err(gv, r_gv) <= 1e-5
    assert rel_err(ga, r_ga) <= 1e-5
    assert torch.equal(res[0][1], res[1][1])


# ---------------------------------------------------------------------------- full-size properties
@pytest.mark.parametrize("vdt,adt", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32)])
def test_full_a2d_shape_properties(vdt, adt):
    """BASELINE config 2 shape (16 frames x 5100 tokens).  Size-independent checks:
    * value == 1 and every sample strictly inside the map  =>  output == sum of weights == 1;
    * then sum over pixels of grad_value[n, :, m, c] == sum over queries of grad_output[n, :, m, c];
    * linearity in value: f(2 v1 - v2) == 2 f(v1) - f(v2);
    * frames are independent: frame 5 alone reproduces frame 5 of the batch bit for bit."""
    x = make_inputs(N=16, dist="encoder", seed=7)
    sh, ls = x.spatial_shapes.to(DEV), x.level_start_index.to(DEV)
    wh = x.spatial_shapes.flip(-1).float().view(1, 1, 1, -1, 1, 2)
    loc = ((x.sampling_locations * wh).clamp(min=0.75) - 0.0)
    loc = torch.minimum(loc, wh - 0.75) / wh                     # corners all valid
    loc, attn = loc.to(DEV, adt).contiguous(), x.attention_weights.to(DEV, adt).contiguous()
    ones = torch.ones_like(x.value).to(DEV, vdt)
    go = x.grad_output.to(DEV, vdt)
    out = msda_ext.ms_deform_attn_forward(ones, sh, ls, loc, attn, 64)
    tol = 1e-5 if vdt == torch.float32 else 1e-2
    assert float((out.float() - 1).abs().max()) <= tol
    gv, gl, ga = msda_ext.ms_deform_attn_backward(ones, sh, ls, loc, attn, go, 64)
    lhs = gv.float().sum(1)                                      # (N, M, D)
    rhs = go.float().view(16, -1, 8, 32).sum(1)
    assert float((lhs - rhs).abs().max()) <= (1e-2 if vdt == torch.float32 else 1.0)
    assert float(gl.float().abs().max()) <= (1e-3 if vdt == torch.float32 else 0.5)  # constant image: no slope

    v1 = x.value.to(DEV, vdt)
    v2 = torch.roll(x.value, 1, 1).to(DEV, vdt)
    loc_r, attn_r = x.sampling_locations.to(DEV, adt), x.attention_weights.to(DEV, adt)
    f1 = msda_ext.ms_deform_attn_forward(v1, sh, ls, loc_r, attn_r, 64).float()
    f2 = msda_ext.ms_deform_attn_forward(v2, sh, ls, loc_r, attn_r, 64).float()
    f3 = msda_ext.ms_deform_attn_forward((2 * v1.float() - v2.float()).to(vdt), sh, ls, loc_r, attn_r, 64).float()
    assert float((f3 - (2 * f1 - f2)).abs().max()) <= (2e-5 if vdt == torch.float32 else 8e-2)

    one = msda_ext.ms_deform_attn_forward(v1[5:6].contiguous(), sh, ls, loc_r[5:6].contiguous(),
                                          attn_r[5:6].contiguous(), 64)
    assert torch.equal(one[0], msda_ext.ms_deform_attn_forward(v1, sh, ls, loc_r, attn_r, 64)[5])
    go1 = go[5:6].contiguous()
    g_one = msda_ext.ms_deform_attn_backward(v1[5:6].contiguous(), sh, ls, loc_r[5:6].contiguous(),
                                             attn_r[5:6].contiguous(), go1, 64)
    g_all = msda_ext.ms_deform_attn_backward(v1, sh, ls, loc_r, attn_r, go, 64)
    for a, b in zip(g_one, g_all):
        assert torch.equal(a[0], b[5])


def test_full_a2d_shape_vs_grid_sample_port_on_gpu():
    """16 frames at once against the reference's own formulation (grid_sample) evaluated by torch
    on the same GPU in fp32 -- the N = 16 counterpart of the fp64 CPU check above."""
    x = make_inputs(N=16, dist="encoder", seed=8).to(DEV)
    out = msda_ext.ms_deform_attn_forward(x.value, x.spatial_shapes, x.level_start_index,
                                          x.sampling_locations, x.attention_weights, 64)
    gv, gl, ga = msda_ext.ms_deform_attn_backward(x.value, x.spatial_shapes, x.level_start_index,
                                                  x.sampling_locations, x.attention_weights, x.grad_output, 64)
    ref = O.grid_sample_port_grads(x.value, x.spatial_shapes.cpu(), x.sampling_locations, x.attention_weights,
                                   x.grad_output)
    keep = off_lattice(x.sampling_locations.cpu().numpy(), x.spatial_shapes.cpu().numpy(), 1e-3)
    assert rel_err(out, ref[0].double().cpu().numpy()) <= 1e-5
    assert rel_err(gv, ref[1].double().cpu().numpy()) <= 2e-5     # torch's own scatter is fp32-atomic
    assert rel_err(ga, ref[3].double().cpu().numpy()) <= 1e-5
    assert rel_err(gl, ref[2].double().cpu().numpy(), keep) <= 5e-5


# ---------------------------------------------------------------------------- autograd boundary
def test_autograd_function_and_gradcheck():
    """models/ops/test.py:63-78: fp64 gradcheck of the op for several channel counts."""
    shapes = torch.tensor([(6, 4), (3, 2)], dtype=torch.long, device=DEV)
    lsi = torch.tensor([0, 24], dtype=torch.long, device=DEV)
    g = torch.Generator().manual_seed(3)
    for D in (30, 32, 64, 71):
        value = (torch.rand(1, 30, 2, D, generator=g) * 0.01).double().to(DEV).requires_grad_(True)
        loc = torch.rand(1, 2, 2, 2, 2, 2, generator=g).double().to(DEV).requires_grad_(True)
        attn = torch.rand(1, 2, 2, 2, 2, generator=g) + 1e-5
        attn = (attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)).double().to(DEV).requires_grad_(True)
        assert torch.autograd.gradcheck(MSDeformAttnFunction.apply, (value, shapes, lsi, loc, attn, 2))


def test_error_behaviour():
    x = make_inputs(N=3, dist="decoder", Lq=4, shapes=[(4, 4)], M=2, D=32, seed=9)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        msda_ext.ms_deform_attn_forward(x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations,
                                        x.attention_weights, 64)
    d = x.to(DEV)
    with pytest.raises(RuntimeError, match="contiguous"):
        msda_ext.ms_deform_attn_forward(d.value.transpose(1, 2), d.spatial_shapes, d.level_start_index,
                                        d.sampling_locations, d.attention_weights, 64)
    with pytest.raises(RuntimeError, match="must divide im2col_step"):   # ms_deform_attn_cuda.cu:52
        msda_ext.ms_deform_attn_forward(d.value, d.spatial_shapes, d.level_start_index, d.sampling_locations,
                                        d.attention_weights, 2)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        msda_ext.ms_deform_attn_forward(d.value, x.spatial_shapes, d.level_start_index, d.sampling_locations,
                                        d.attention_weights, 64)
    out = msda_ext.ms_deform_attn_forward(d.value, d.spatial_shapes, d.level_start_index, d.sampling_locations,
                                          d.attention_weights, 3)
    assert out.shape == (3, 4, 64) and msda_ext.last_launch_count() == 1


def test_cuda_graph_capture_and_replay():
    """Every launch of a forward + backward is asynchronous on the caller's stream with no hidden
    synchronisation, so the pair can be captured once and replayed."""
    x = make_inputs(N=2, dist="encoder", shapes=[(12, 20), (6, 10), (3, 5), (2, 3)], seed=11).to(DEV)
    args = (x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights)
    eager_out, index = msda_ext.ms_deform_attn_forward(*args, 64, want_index=True)
    eager = msda_ext.ms_deform_attn_backward(*args, x.grad_output, 64, index=index)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            out, idx = msda_ext.ms_deform_attn_forward(*args, 64, want_index=True)
            grads = msda_ext.ms_deform_attn_backward(*args, x.grad_output, 64, index=idx)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager_out)
    for a, b in zip(grads, eager):
        assert torch.equal(a, b)


# ---------------------------------------------------------------------------- odd geometries
def _check_against_oracle(x, vdt=torch.float32, adt=torch.float32, tol=1e-5, flags=0):
    out, gv, gl, ga, (v, lo, at, go) = run_op(x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations,
                                               x.attention_weights, x.grad_output, vdt, adt, flags=flags)
    r_out, r_gv, r_gl, r_ga = oracle_f64(v, x.spatial_shapes, x.level_start_index, lo, at, go)
    keep = off_lattice(lo.double().cpu().numpy(), x.spatial_shapes.numpy(), 1e-3)
    assert rel_err(out, r_out) <= tol
    assert rel_err(gv, r_gv) <= tol
    assert rel_err(ga, r_ga) <= tol
    assert rel_err(gl, r_gl, keep) <= 2 * tol


def test_level_start_index_with_gaps():
    """value rows that belong to no level (level_start_index leaves gaps) take part in nothing and must
    get an exactly zero gradient."""
    x = make_inputs(N=2, dist="uniform", shapes=[(5, 6), (3, 4)], M=4, D=32, Lq=40, seed=31)
    S2 = 30 + 7 + 12 + 5                                  # 7 unused rows between the levels, 5 at the end
    value = torch.randn(2, S2, 4, 32, generator=torch.Generator().manual_seed(1))
    x.value = value
    x.level_start_index = torch.tensor([0, 37], dtype=torch.long)
    out, gv, gl, ga, (v, lo, at, go) = run_op(x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations,
                                               x.attention_weights, x.grad_output, torch.float32, torch.float32)
    r_out, r_gv, r_gl, r_ga = oracle_f64(v, x.spatial_shapes, x.level_start_index, lo, at, go)
    assert rel_err(out, r_out) <= 1e-5 and rel_err(gv, r_gv) <= 1e-5 and rel_err(ga, r_ga) <= 1e-5
    assert float(gv[:, 30:37].abs().max()) == 0.0 and float(gv[:, 49:].abs().max()) == 0.0


@pytest.mark.parametrize("kw", [
    dict(N=1, shapes=[(1, 1)], M=1, D=32, Lq=3, P=4, dist="uniform"),                       # a single pixel
    dict(N=1, shapes=[(1, 37), (23, 1)], M=2, D=32, Lq=50, P=4, dist="uniform"),            # 1-pixel-wide maps
    dict(N=3, shapes=[(2, 2)] * 16, M=2, D=32, Lq=9, P=4, dist="uniform"),                  # 16 levels
    dict(N=1, shapes=[(40, 64), (20, 32)], M=1, D=64, Lq=700, P=8, dist="uniform"),         # P = 8, D = 64
    dict(N=2, shapes=[(9, 11)], M=3, D=16, Lq=130, P=4, dist="uniform"),                    # D = 16, one level
    dict(N=1, shapes=[(6, 7), (3, 4)], M=2, D=32, Lq=257, P=16, dist="uniform"),            # P = 16 (generic path)
])
def test_odd_geometries_fp32(kw):
    _check_against_oracle(make_inputs(seed=41, **kw))


def test_everything_out_of_range_and_nan_locations():
    """No accepted sample at all: outputs and every gradient are exactly zero (cuh:288,365-374), also for
    NaN locations, which no comparison accepts."""
    x = make_inputs(N=1, dist="uniform", shapes=[(6, 8), (3, 4)], M=2, D=32, Lq=40, seed=51)
    loc = x.sampling_locations * 0 + 3.0
    loc[:, ::2] = float("nan")
    out, gv, gl, ga, _ = run_op(x.value, x.spatial_shapes, x.level_start_index, loc, x.attention_weights,
                                x.grad_output, torch.float32, torch.float32)
    for t in (out, gv, gl, ga):
        assert float(t.abs().max()) == 0.0


def test_non_finite_values_stay_local():
    """An Inf in one value row reaches only the outputs whose accepted corners touch it, exactly as in the
    reference (rejected samples and corners outside the map read nothing)."""
    x = make_inputs(N=1, dist="uniform", shapes=[(8, 8)], M=1, D=32, Lq=64, P=4, seed=61)
    value = x.value.clone()
    value[0, 27] = float("inf")                            # pixel (3, 3)
    dev = x.to(DEV)
    out = msda_ext.ms_deform_attn_forward(value.to(DEV), dev.spatial_shapes, dev.level_start_index,
                                          dev.sampling_locations, dev.attention_weights, 64)
    ref = O.forward_c(value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights,
                      dtype=np.float32)
    got = out.cpu().numpy()
    assert np.array_equal(np.isfinite(got), np.isfinite(ref.reshape(got.shape)))


def test_large_query_count_and_frames():
    """Config 4 / 5 sizes: 36 frames, and 20 k tokens per frame (one frame), forward + backward vs the
    grid_sample formulation on the GPU."""
    for kw in (dict(N=36, dist="decoder", Lq=5), dict(N=1, dist="encoder", shapes=[(96, 160), (48, 80), (24, 40), (12, 20)])):
        x = make_inputs(seed=71, **kw).to(DEV)
        out = msda_ext.ms_deform_attn_forward(x.value, x.spatial_shapes, x.level_start_index,
                                              x.sampling_locations, x.attention_weights, 64)
        gv, gl, ga = msda_ext.ms_deform_attn_backward(x.value, x.spatial_shapes, x.level_start_index,
                                                      x.sampling_locations, x.attention_weights, x.grad_output, 64)
        ref = O.grid_sample_port_grads(x.value, x.spatial_shapes.cpu(), x.sampling_locations, x.attention_weights,
                                       x.grad_output)
        keep = off_lattice(x.sampling_locations.cpu().numpy(), x.spatial_shapes.cpu().numpy(), 1e-3)
        assert rel_err(out, ref[0].double().cpu().numpy()) <= 1e-5
        assert rel_err(gv, ref[1].double().cpu().numpy()) <= 2e-5
        assert rel_err(ga, ref[3].double().cpu().numpy()) <= 1e-5
        assert rel_err(gl, ref[2].double().cpu().numpy(), keep) <= 5e-5


# ---------------------------------------------------------------------------- module level
@pytest.mark.parametrize("amp", [False, True])
def test_module_fused_prologue_matches_unfused(amp):
    """MSDeformAttn with softmax + location arithmetic fused into the forward kernel against the same module
    running the reference's elementwise sequence (ms_deform_attn.py:99-106) around the plain op: outputs,
    returned locations / weights and every parameter and input gradient."""
    from neurips2023_soc_b200 import MSDeformAttn
    torch.manual_seed(0)
    shapes_l = [(12, 20), (6, 10), (3, 5), (2, 3)]
    S = sum(h * w for h, w in shapes_l)
    shapes = torch.tensor(shapes_l, dtype=torch.long, device=DEV)
    lsi = torch.tensor([0, 240, 300, 315], dtype=torch.long, device=DEV)
    mod = MSDeformAttn(256, 4, 8, 4).to(DEV)
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.05)
        mod.attention_weights.weight.normal_(0, 0.2)
    N, Lq = 2, 37
    query = torch.randn(N, Lq, 256, device=DEV, requires_grad=True)
    src = torch.randn(N, S, 256, device=DEV, requires_grad=True)
    ref = torch.rand(N, Lq, 4, 2, device=DEV, requires_grad=True)
    pad = torch.zeros(N, S, dtype=torch.bool, device=DEV)
    pad[1, -9:] = True
    g = torch.randn(N, Lq, 256, device=DEV)

    def run(fused):
        mod.fused_prologue = fused
        for t in (query, src, ref):
            t.grad = None
        mod.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            if fused:
                out, loc, w = mod(query, ref, src, shapes, lsi, pad)
            else:
                # the reference's sequence with the elementwise part evaluated in fp32 from the same
                # (possibly bf16) projections -- under autocast the stock sequence rounds offset / (W, H)
                # to bf16 first, which moves samples across pixel borders and makes grad_loc incomparable
                value = mod.value_proj(src).masked_fill(pad[..., None], 0.0).view(N, S, 8, 32)
                off = mod.sampling_offsets(query).view(N, Lq, 8, 4, 4, 2).float()
                w = torch.softmax(mod.attention_weights(query).view(N, Lq, 8, 16).float(), -1).view(N, Lq, 8, 4, 4)
                loc = ref[:, :, None, :, None, :] + off / shapes.flip(-1)[None, None, None, :, None, :]
                out = mod.output_proj(MSDeformAttnFunction.apply(value, shapes, lsi, loc, w, 64))
        _lib.profile_enable(True)              # process-wide: sees the kernels autograd's thread launches
        (out.float() * g).sum().backward()
        torch.cuda.synchronize()
        kernels = [name for name, _ in _lib.profile_read()]
        _lib.profile_enable(False)
        grads = [p.grad.clone() for p in mod.parameters()] + [query.grad.clone(), src.grad.clone(), ref.grad.clone()]
        return out.detach().float(), loc.detach().float(), w.detach().float(), grads, kernels

    a, b = run(True), run(False)
    # nobody differentiates through the returned locations / weights here, so the fused module must take the
    # in-kernel chain rule (msda_backward_fused), the unfused one the plain sample-gradient kernel
    assert "msda_bwd_sample_tile_kernel<chain>" in a[4], a[4]
    # (37 queries per frame: the unfused backward of such a call is the one-launch kernel)
    assert "msda_bwd_direct_kernel" in b[4] and "msda_bwd_sample_tile_kernel<chain>" not in b[4], b[4]
    tol = 3e-2 if amp else 2e-5
    assert a[1].dtype == torch.float32 and a[1].shape == (N, Lq, 8, 4, 4, 2)
    assert float((a[1] - b[1]).abs().max()) <= 1e-6                          # sampling locations
    assert float((a[2] - b[2]).abs().max()) <= 1e-6                          # attention weights
    assert float((a[0] - b[0]).abs().max()) <= tol * max(1.0, float(b[0].abs().max()))
    for x, y in zip(a[3], b[3]):
        assert float((x.float() - y.float()).abs().max()) <= tol * max(1.0, float(y.float().abs().max()))


def test_bf16_lane_layout_switch():
    """MSDA_FLAG_BF16_VEC4 (64-byte bf16 rows on 8 lanes x 64 bit instead of 4 lanes x 128 bit) is an A/B
    switch: same arithmetic per row, so results agree to the last bit in the forward and to bf16 rounding in
    the backward (the cross-lane reduction tree differs)."""
    x = make_inputs(N=2, dist="encoder", shapes=[(12, 20), (6, 10), (3, 5), (2, 3)], seed=81)
    a = run_op(x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights,
               x.grad_output, torch.bfloat16, torch.float32)
    b = run_op(x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights,
               x.grad_output, torch.bfloat16, torch.float32, flags=_lib.FLAG_BF16_VEC4)
    assert torch.equal(a[0], b[0])
    for i in (1, 2, 3):
        assert rel_err(a[i], b[i].double().cpu().numpy()) <= 2e-2


@pytest.mark.parametrize("frames_per_chunk,ramp", [(1, False), (2, False), (2, True), (3, False), (8, True)])
def test_host_frame_pipeline_matches_one_call(frames_per_chunk, ramp):
    """The host entry point (chunks of frames streamed over PCIe on three streams) returns the bits of
    one call over the whole batch, for ragged last chunks and when called again on the same state."""
    from neurips2023_soc_b200.host_frames import HostFramePipeline
    x = make_inputs(N=7, shapes=[(12, 20), (6, 10), (3, 5), (2, 3)], dist="encoder", seed=11)
    host = x.to("cpu", torch.bfloat16, torch.float32)
    pins = [t.pin_memory() for t in (host.value, host.sampling_locations, host.attention_weights, host.grad_output)]
    d = host.to(DEV)
    a = (d.value, d.spatial_shapes, d.level_start_index, d.sampling_locations, d.attention_weights)
    out, index = msda_ext.ms_deform_attn_forward(*a, 64, want_index=True)
    ref = [out] + msda_ext.ms_deform_attn_backward(*a, d.grad_output, 64, index=index)
    pipe = HostFramePipeline(DEV, frames_per_chunk=frames_per_chunk, ramp=ramp)
    for _ in range(2):
        res = pipe.forward_backward(pins[0], host.spatial_shapes, host.level_start_index, pins[1], pins[2], pins[3])
        torch.cuda.current_stream().synchronize()
        assert pipe.launches > 0
        for got, want in zip(res, ref):
            assert got.is_pinned() and torch.equal(got, want.cpu())
    with pytest.raises(RuntimeError, match="pinned"):
        pipe.forward_backward(host.value, host.spatial_shapes, host.level_start_index, pins[1], pins[2], pins[3])


def test_host_frame_pipeline_as_cuda_graph():
    """graph=True: the step is captured once per set of host buffers and replayed; every call (the capturing one,
    replays, replays after the inputs changed in place, another set of buffers) returns the bits of one call."""
    from neurips2023_soc_b200.host_frames import HostFramePipeline
    x = make_inputs(N=7, shapes=[(12, 20), (6, 10), (3, 5), (2, 3)], dist="encoder", seed=12)
    host = x.to("cpu", torch.bfloat16, torch.float32)
    pins = [t.pin_memory() for t in (host.value, host.sampling_locations, host.attention_weights, host.grad_output)]

    def reference():
        d = [p.to(DEV) for p in pins]
        sh, ls = host.spatial_shapes.to(DEV), host.level_start_index.to(DEV)
        out, index = msda_ext.ms_deform_attn_forward(d[0], sh, ls, d[1], d[2], 64, want_index=True)
        return [out] + msda_ext.ms_deform_attn_backward(d[0], sh, ls, d[1], d[2], d[3], 64, index=index)

    pipe = HostFramePipeline(DEV, frames_per_chunk=2, graph=True)
    results = None
    for it in range(4):
        if it == 2:                                    # new contents in the same host buffers: the replay must see them
            pins[0].mul_(0.5)
            pins[3].add_(0.25)
        ref = reference()
        res = pipe.forward_backward(pins[0], host.spatial_shapes, host.level_start_index, pins[1], pins[2], pins[3],
                                    results=results)
        torch.cuda.current_stream().synchronize()
        results = res
        for got, want in zip(res, ref):
            assert torch.equal(got, want.cpu())
    assert len(pipe._graphs) == 1
    other = [p.clone().pin_memory() for p in pins]     # other buffers: their own graph
    res2 = pipe.forward_backward(other[0], host.spatial_shapes, host.level_start_index, other[1], other[2], other[3])
    torch.cuda.current_stream().synchronize()
    for got, want in zip(res2, reference()):
        assert torch.equal(got, want.cpu())
    assert len(pipe._graphs) == 2


@pytest.mark.parametrize("vdt,adt", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32),
                                      (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("kw", [dict(N=3, dist="decoder", Lq=20), dict(N=2, dist="decoder", Lq=5),
                                dict(N=2, dist="decoder", Lq=128), dict(N=1, dist="decoder", Lq=1),
                                dict(N=2, dist="uniform", Lq=100),
                                dict(N=2, dist="uniform", shapes=[(12, 20), (6, 10), (3, 5), (2, 3)], Lq=64, P=8),
                                dict(N=1, dist="uniform", shapes=[(9, 11), (4, 5)], M=3, D=64, Lq=60, P=8),
                                dict(N=2, dist="uniform", shapes=[(7, 9)], M=2, D=16, Lq=33, P=4)])
def test_grad_value_gathers_agree(kw, vdt, adt):
    """grad_value has two routes: the inverse-index pipeline (count / scan / fill / sort / walk) and, for calls
    with few queries per frame (4 * Lq * P <= 2048: decoder cross-attention), a direct gather that sorts one
    (frame, head, level)'s contributions in shared memory.  Forced either way, both must match the oracle and
    leave grad_loc / grad_attn untouched (bit for bit with the three-launch sequence, which shares the sample-gradient
    kernel with the index pipeline; to rounding with the one-launch kernel, which reduces over lanes in another order);
    the direct one must be bit-reproducible and is the default here."""
    if vdt == torch.float32 and adt != torch.float32:
        pytest.skip("fp32 values take fp32 locations")
    x = make_inputs(seed=17, **kw)
    a = (x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights, x.grad_output)
    dense = run_op(*a, vdt, adt, flags=_lib.FLAG_WALK_DENSE)
    direct = run_op(*a, vdt, adt)
    direct2 = run_op(*a, vdt, adt)
    split = run_op(*a, vdt, adt, flags=_lib.FLAG_DIRECT_SPLIT)
    v, lo, at, go = dense[4]
    r_gv = oracle_f64(v, x.spatial_shapes, x.level_start_index, lo, at, go)[1]
    tol = TOL[vdt]
    assert rel_err(dense[1], r_gv) <= tol and rel_err(direct[1], r_gv) <= tol
    assert torch.equal(direct[1], direct2[1]) and torch.equal(direct[1], split[1])
    for i in (0, 2, 3):
        assert torch.equal(dense[i], split[i])
        assert torch.equal(direct[i], direct2[i])
        assert rel_err(direct[i], dense[i].double().cpu().numpy()) <= tol
    N, S, M, D = x.value.shape
    Lq, L, P = x.sampling_locations.shape[1], x.sampling_locations.shape[3], x.sampling_locations.shape[4]
    assert _lib.load().msda_index_bytes(N, S, M, D, L, Lq, P) == 0       # no index handoff for these shapes
    assert _lib.load().msda_index_bytes(N, S, M, D, L, 600, P) > 0


@pytest.mark.parametrize("N,Lq", [(16, 20), (36, 5), (36, 20), (16, 128)])
def test_direct_gather_at_decoder_batch_sizes(N, Lq):
    """The SOC decoder calls (A2D: 16 frames x 20 queries; Ref-YouTube-VOS: 36 frames x 5 / 20 queries) with every
    SM busy: the direct gather against the index pipeline (different summation order, so to tolerance) and
    against itself (bit for bit, 3 runs)."""
    x = make_inputs(N=N, Lq=Lq, dist="decoder", seed=29)
    a = (x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights, x.grad_output)
    dense = run_op(*a, torch.float32, torch.float32, flags=_lib.FLAG_WALK_DENSE)
    runs = [run_op(*a, torch.float32, torch.float32) for _ in range(3)]
    assert rel_err(runs[0][1], dense[1].double().cpu().numpy()) <= 1e-5
    assert torch.equal(runs[0][1], runs[1][1]) and torch.equal(runs[0][1], runs[2][1])
    x1 = make_inputs(N=1, Lq=Lq, dist="decoder", seed=29)
    v, lo, at, go = (t[:1] for t in dense[4])
    r_gv = oracle_f64(v, x.spatial_shapes, x.level_start_index, lo, at, go)[1]
    assert rel_err(runs[0][1][:1], r_gv) <= 1e-5


def test_direct_gather_is_frame_independent():
    """The gather is chosen from per-frame quantities only: a frame's gradients do not depend on its batch."""
    x = make_inputs(N=4, dist="decoder", Lq=20, seed=23)
    a = (x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights, x.grad_output)
    full = run_op(*a, torch.float32, torch.float32)
    one = run_op(x.value[2:3], x.spatial_shapes, x.level_start_index, x.sampling_locations[2:3],
                 x.attention_weights[2:3], x.grad_output[2:3], torch.float32, torch.float32)
    for i in range(4):
        assert torch.equal(full[i][2:3], one[i])


@pytest.mark.parametrize("vdt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("kw", [dict(N=2, dist="encoder", shapes=[(12, 20), (6, 10), (3, 5), (2, 3)]),
                                dict(N=2, dist="uniform", shapes=[(12, 20), (6, 10), (3, 5), (2, 3)], Lq=300),
                                dict(N=3, dist="decoder", Lq=20),
                                dict(N=1, dist="uniform", shapes=[(9, 11), (4, 5)], M=3, D=64, Lq=70, P=8)])
def test_backward_fused_applies_the_prologue_chain_rule(kw, vdt):
    """msda_backward_fused against msda_backward_indexed followed by the chain rule of the module's prologue in
    torch: grad_offsets = grad_loc / (W, H), grad_logits = softmax backward over the L*P weights of a (query,
    head) -- including samples the op rejects (zero grad_attn, yet their logit takes -a * sum).  grad_value is
    the same kernels' output: bit for bit."""
    x = make_inputs(seed=53, **kw)
    d = x.to(DEV, vdt, torch.float32)
    a = (d.value, d.spatial_shapes, d.level_start_index, d.sampling_locations, d.attention_weights)
    _, index = msda_ext.ms_deform_attn_forward(*a, 64, want_index=True)
    gv, gl, ga = msda_ext.ms_deform_attn_backward(*a, d.grad_output, 64, index=index)
    _, index = msda_ext.ms_deform_attn_forward(*a, 64, want_index=True)
    fv, foff, flog = msda_ext.ms_deform_attn_backward_fused(*a, d.grad_output, 64, index=index)
    wh = d.spatial_shapes.flip(-1).float()[None, None, None, :, None, :]
    attn = d.attention_weights
    want_off = gl / wh
    want_log = attn * (ga - (ga * attn).sum(dim=(-1, -2), keepdim=True))
    assert torch.equal(fv, gv)
    scale = lambda t: max(1.0, float(t.abs().max()))   # noqa: E731
    assert float((foff - want_off).abs().max()) <= 1e-5 * scale(want_off)
    assert float((flog - want_log).abs().max()) <= 1e-5 * scale(want_log)
    assert float(flog.abs().max()) > 0


def test_backward_fused_rejects_what_it_has_no_kernel_for():
    x = make_inputs(N=1, shapes=[(6, 7), (3, 4)], M=2, D=20, Lq=9, P=4, dist="uniform", seed=3).to(DEV)
    with pytest.raises(RuntimeError, match="fused prologue"):
        msda_ext.ms_deform_attn_backward_fused(x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations,
                                               x.attention_weights, x.grad_output, 64)


# ---------------------------------------------------------------------------- grad_value bin kernel (part B, A/B arm)
@pytest.mark.parametrize("vdt,adt", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32)])
@pytest.mark.parametrize("kw", [
    dict(N=2, dist="encoder"),                                                               # the A2D pyramid
    dict(N=1, dist="uniform"),                                                               # adversarial locations
    dict(N=1, dist="encoder", shapes=[(50, 77), (25, 39), (13, 20), (7, 10)]),               # tiles that do not divide
    dict(N=2, dist="uniform", shapes=[(9, 11)], M=3, D=64, Lq=700, P=8),                     # D = 64, P = 8, one level
    dict(N=1, dist="uniform", shapes=[(3, 4), (1, 2)], M=2, D=32, Lq=3000, P=4),             # dense tiny maps: long bins
])
def test_grad_value_bin_kernel_vs_sort_and_walk(kw, vdt, adt):
    """The one-kernel bin pass (MSDA_FLAG_BIN_KERNEL) and the rank-sort + row-walker pair (default) sum the same
    terms in different orders: they agree to rounding, both agree with the fp64 oracle, and part A's results are
    the same bits either way."""
    x = make_inputs(seed=81, **kw)
    a = (x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights, x.grad_output)
    new = run_op(*a, vdt, adt, flags=_lib.FLAG_BIN_KERNEL)
    old = run_op(*a, vdt, adt)
    v, lo, at, go = new[4]
    r_gv = oracle_f64(v, x.spatial_shapes, x.level_start_index, lo, at, go)[1]
    tol = TOL[vdt]
    assert rel_err(new[1], r_gv) <= tol
    assert rel_err(old[1], r_gv) <= tol
    assert torch.equal(new[0], old[0]) and torch.equal(new[2], old[2]) and torch.equal(new[3], old[3])


@pytest.mark.parametrize("vdt,adt", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32)])
def test_unordered_flag_skips_the_sort_only(vdt, adt):
    """MSDA_FLAG_UNORDERED (opt-in) leaves the index entries in arrival order: grad_value is the same sum in another
    order (within the dtype's bound of the fp64 oracle, like the reference's atomics), the sample gradients and the
    output are the same bits, and no sort kernel is launched."""
    x = make_inputs(N=2, dist="encoder", seed=85)
    a = (x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights, x.grad_output)
    _lib.profile_enable(True)
    new = run_op(*a, vdt, adt, flags=_lib.FLAG_UNORDERED)
    torch.cuda.synchronize()
    names = [n for n, _ in _lib.profile_read()]
    _lib.profile_enable(False)
    assert "msda_grad_value_walk_kernel" in names and not any("sort" in n for n in names), names
    old = run_op(*a, vdt, adt)
    v, lo, at, go = new[4]
    r_gv = oracle_f64(v, x.spatial_shapes, x.level_start_index, lo, at, go)[1]
    assert rel_err(new[1], r_gv) <= TOL[vdt]
    assert torch.equal(new[0], old[0]) and torch.equal(new[2], old[2]) and torch.equal(new[3], old[3])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_c_abi_runs_on_the_device_that_owns_the_pointers(monkeypatch):
    """the ABI names its device through the pointers.  With device 0 current and no device context around
    the call (torch.cuda.device patched out of the shim), tensors on cuda:1 are processed on cuda:1 and device 0 is
    current again afterwards."""
    x = make_inputs(N=1, dist="encoder", seed=86).to("cuda:1", torch.float32, torch.float32)
    a = (x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights)
    ref_out = msda_ext.ms_deform_attn_forward(*a, 64)
    ref_g = msda_ext.ms_deform_attn_backward(*a, x.grad_output, 64)
    torch.cuda.synchronize(1)

    class NoDeviceContext:
        def __init__(self, *args):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    torch.cuda.set_device(0)
    monkeypatch.setattr(torch.cuda, "device", NoDeviceContext)
    out = msda_ext.ms_deform_attn_forward(*a, 64)          # stream: device 0's current stream = the default stream
    g = msda_ext.ms_deform_attn_backward(*a, x.grad_output, 64)
    monkeypatch.undo()
    assert torch.cuda.current_device() == 0
    torch.cuda.synchronize(1)
    assert torch.equal(out, ref_out)
    for u, w in zip(g, ref_g):
        assert torch.equal(u, w)


@pytest.mark.parametrize("flags", [0, _lib.FLAG_BIN_KERNEL])
def test_grad_value_bf16_wide_rows(flags):
    """bf16 rows of 128 and 256 bytes (D = 64, 128): 8 and 16 lanes x 128 bit per row."""
    for D in (64, 128):
        x = make_inputs(N=1, dist="uniform", shapes=[(10, 12), (5, 6)], M=2, D=D, Lq=400, P=4, seed=82)
        _check_against_oracle(x, torch.bfloat16, torch.float32, tol=2e-2, flags=flags)


@pytest.mark.parametrize("flags", [0, _lib.FLAG_BIN_KERNEL])
@pytest.mark.parametrize("Lq", [12000, 40000, 70000])
def test_grad_value_oversized_sub_bins(Lq, flags):
    """Every query samples the same spot of a 2 x 2 map: one bin holds Lq * P entries in 256 sub-bins of 187
    (sorted in place beforehand, kept in order by the rank step), 625 or 1094 entries (larger than one staging
    chunk: sliced).  Checked against the fp64 oracle, and twice for bit reproducibility."""
    x = make_inputs(N=1, dist="uniform", shapes=[(2, 2)], M=1, D=32, Lq=Lq, P=4, seed=83)
    loc = x.sampling_locations * 0.0 + torch.tensor([0.52, 0.47])
    loc[:, ::97] += 0.2                                       # a few elsewhere
    res = [run_op(x.value, x.spatial_shapes, x.level_start_index, loc, x.attention_weights, x.grad_output,
                  torch.float32, torch.float32, flags=flags) for _ in range(2)]
    out, gv, gl, ga, (v, lo, at, go) = res[0]
    r_out, r_gv, r_gl, r_ga = oracle_f64(v, x.spatial_shapes, x.level_start_index, lo, at, go)
    assert rel_err(gv, r_gv) <= 1e-5
    assert rel_err(ga, r_ga) <= 1e-5
    assert torch.equal(res[0][1], res[1][1])


@pytest.mark.parametrize("flags", [0, _lib.FLAG_BIN_KERNEL])
def test_grad_value_is_frame_independent(flags):
    """Tile sizes, round cuts and the sharing of dense bins depend on per-frame quantities only: a frame alone
    reproduces its slice of the batch bit for bit (fp32 and bf16, both location distributions)."""
    for dist, vdt in (("encoder", torch.float32), ("uniform", torch.bfloat16)):
        x = make_inputs(N=3, dist=dist, seed=84).to(DEV, vdt, torch.float32)
        a = (x.spatial_shapes, x.level_start_index)
        g_all = msda_ext.ms_deform_attn_backward(x.value, *a, x.sampling_locations, x.attention_weights, x.grad_output, 64,
                                                 flags=flags)
        g_one = msda_ext.ms_deform_attn_backward(x.value[1:2].contiguous(), *a, x.sampling_locations[1:2].contiguous(),
                                                 x.attention_weights[1:2].contiguous(), x.grad_output[1:2].contiguous(), 64,
                                                 flags=flags)
        for u, w in zip(g_one, g_all):
            assert torch.equal(u[0], w[1])


# ---------------------------------------------------------------------------- the benchmarked configuration itself
def test_bench_configuration_vs_oracle_all_16_frames():
    """bench.py's own step -- 16 frames x 5100 tokens, bf16 values, fp32 locations / weights, the index handed from
    the forward to the backward, buffers passed in -- against the fp64 C oracle evaluated on the bf16-rounded
    inputs, every frame (2e-2, north_star's bf16 bound)."""
    x = make_inputs(N=16, dist="encoder", seed=0)
    out, gv, gl, ga, (v, lo, at, go) = run_op(x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations,
                                               x.attention_weights, x.grad_output, torch.bfloat16, torch.float32)
    r_out, r_gv, r_gl, r_ga = oracle_f64(v, x.spatial_shapes, x.level_start_index, lo, at, go)
    keep = off_lattice(lo.double().cpu().numpy(), x.spatial_shapes.numpy(), 1e-3)
    worst = {}
    for n in range(16):          # frame by frame: no frame hides behind another's larger values
        worst["out"] = max(worst.get("out", 0.0), rel_err(out[n], r_out[n]))
        worst["grad_value"] = max(worst.get("grad_value", 0.0), rel_err(gv[n], r_gv[n]))
        worst["grad_attn"] = max(worst.get("grad_attn", 0.0), rel_err(ga[n], r_ga[n]))
        worst["grad_loc"] = max(worst.get("grad_loc", 0.0), rel_err(gl[n], r_gl[n], keep[n]))
    assert max(worst.values()) <= 2e-2, worst


@pytest.mark.parametrize("D", [1025, 2048, 3096])
def test_large_channel_counts_of_the_reference_test(D):
    """models/ops/test.py:85 walks through D = 1025, 2048, 3096 to reach the reference's global-memory and
    multi-block backward variants (cuh:731-920); here they take the any-D kernels.  fp64 against the C oracle."""
    x = make_inputs(N=1, dist="uniform", shapes=[(6, 4), (3, 2)], M=2, D=D, Lq=2, P=2, seed=91, value_scale=0.01)
    _check_against_oracle(x, torch.float64, torch.float64, tol=1e-11)
    _check_against_oracle(x, torch.float32, torch.float32, tol=1e-5)


# ---------------------------------------------------------------------------- decoder-shaped backward in one launch
@pytest.mark.parametrize("vdt,adt", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32), (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("kw", [dict(N=16, dist="decoder", Lq=20), dict(N=36, dist="decoder", Lq=5),
                                dict(N=2, dist="decoder", Lq=100, P=4), dict(N=3, dist="uniform", Lq=7, P=8, D=64, M=2, shapes=[(9, 11), (5, 6)])])
def test_one_launch_decoder_backward(kw, vdt, adt):
    """Calls with few queries per frame run their whole backward in ONE kernel (zero-fill of grad_value, sample
    gradients, direct gather); MSDA_FLAG_DIRECT_SPLIT keeps the three-launch sequence.  Same grad_value bits (the
    gather is the same), sample gradients equal to rounding, both against the fp64 oracle, and the launch counts."""
    x = make_inputs(seed=95, **kw)
    a = (x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights, x.grad_output)
    one = run_op(*a, vdt, adt)
    n_one = msda_ext.last_launch_count()
    split = run_op(*a, vdt, adt, flags=_lib.FLAG_DIRECT_SPLIT)
    n_split = msda_ext.last_launch_count()
    assert n_one == 1 and n_split == 3
    assert torch.equal(one[1], split[1])
    v, lo, at, go = one[4]
    r_out, r_gv, r_gl, r_ga = oracle_f64(v, x.spatial_shapes, x.level_start_index, lo, at, go)
    keep = off_lattice(lo.double().cpu().numpy(), x.spatial_shapes.numpy(), 1e-3)
    tol = TOL[vdt] if adt == vdt or vdt == torch.float32 else TOL[torch.bfloat16]
    for got in (one, split):
        assert rel_err(got[1], r_gv) <= tol
        assert rel_err(got[3], r_ga) <= tol
        assert rel_err(got[2], r_gl, keep) <= 2 * tol


def test_one_launch_decoder_backward_zero_fills_gaps_and_untouched_rows():
    x = make_inputs(N=2, dist="decoder", shapes=[(5, 6), (3, 4)], M=4, D=32, Lq=6, seed=96)
    S2 = 30 + 7 + 12 + 5                                  # 7 unused rows between the levels, 5 at the end
    value = torch.randn(2, S2, 4, 32, generator=torch.Generator().manual_seed(1))
    lsi = torch.tensor([0, 37], dtype=torch.long)
    dev = [t.to(DEV) for t in (value, x.spatial_shapes, lsi, x.sampling_locations, x.attention_weights, x.grad_output)]
    poison = torch.full_like(dev[0], float("nan"))
    gv, gl, ga = msda_ext.ms_deform_attn_backward(*dev[:5], dev[5], 64, grads=(poison, torch.empty_like(dev[3]), torch.empty_like(dev[4])))
    assert msda_ext.last_launch_count() == 1
    assert torch.isfinite(gv).all()                       # every row was written
    assert float(gv[:, 30:37].abs().max()) == 0.0 and float(gv[:, 49:].abs().max()) == 0.0
    r = oracle_f64(dev[0].cpu(), x.spatial_shapes, lsi, dev[3].cpu(), dev[4].cpu(), dev[5].cpu())
    assert rel_err(gv, r[1]) <= 1e-5


def test_one_launch_decoder_backward_in_a_cuda_graph():
    """The one-launch backward is a cooperative launch (grid-wide barrier between zero-fill and row writes): it must
    survive stream capture and replay like the other kernels."""
    x = make_inputs(N=16, dist="decoder", Lq=20, seed=97).to(DEV)
    a = (x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights)
    grads = (torch.empty_like(x.value), torch.empty_like(x.sampling_locations), torch.empty_like(x.attention_weights))
    ws = torch.empty(max(16, msda_ext.backward_workspace_bytes(x.value, x.sampling_locations)), dtype=torch.uint8, device=DEV)
    eager = [t.clone() for t in msda_ext.ms_deform_attn_backward(*a, x.grad_output, 64, grads=grads, workspace=ws)]
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        msda_ext.ms_deform_attn_backward(*a, x.grad_output, 64, grads=grads, workspace=ws)
    for t in grads:
        t.fill_(float("nan"))
    g.replay()
    torch.cuda.synchronize()
    for u, w in zip(grads, eager):
        assert torch.equal(u, w)
