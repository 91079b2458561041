"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the header
declares, argument validation runs before any CUDA call, and the Python mirror fails loudly
instead of falling back."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

from neurips2023_soc_b200 import MSDeformAttn, MSDeformAttnFunction, _lib, build, msda_ext
from neurips2023_soc_b200.synthetic import algorithmic_bytes, make_inputs, scaled_pyramid

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    build.build_library()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    header = (ROOT / "include" / "msda_b200.h").read_text()
    declared = set(re.findall(r"^\s*(?:int|void|size_t|const char \*)\s*\*?\s*(msda_\w+)\s*\(", header, re.M))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/msda_b200.h but not exported"
    assert lib.msda_version() == int(re.search(r"#define MSDA_VERSION (\d+)", header).group(1))


def test_validation_precedes_cuda(lib):
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)
    assert lib.msda_forward(null, null, null, null, null, null, 1, 1, 1, 32, 1, 1, 1, 0, 0, 64, null) == 1
    assert b"null" in lib.msda_last_error()
    assert lib.msda_forward(one, one, one, one, one, one, 0, 1, 1, 32, 1, 1, 1, 0, 0, 64, null) == 1
    assert lib.msda_forward(one, one, one, one, one, one, 3, 4, 1, 32, 1, 1, 1, 0, 0, 2, null) == 2
    assert b"must divide im2col_step" in lib.msda_last_error()
    assert lib.msda_forward(one, one, one, one, one, one, 2, 4, 1, 32, 1, 1, 1, 0, 1, 64, null) == 1  # aux bf16, value f32
    assert lib.msda_forward(one, one, one, one, one, one, 2, 4, 1, 32, 17, 1, 1, 0, 0, 64, null) == 5  # > 16 levels
    # backward without a workspace
    assert lib.msda_backward(one, one, one, one, one, one, one, one, one, null, 0, 2, 4, 1, 32, 1, 1, 1, 0, 0, 64, null) == 3


def test_workspace_size(lib):
    small = lib.msda_backward_workspace_bytes(1, 5100, 8, 32, 4, 5100, 4, 0, 0)
    big = lib.msda_backward_workspace_bytes(16, 5100, 8, 32, 4, 5100, 4, 0, 0)
    assert 0 < small < big
    samples = 16 * 5100 * 8 * 16
    assert big >= samples * 16                            # one 16-byte entry per sample
    assert big < samples * 16 + (64 << 20)                # plus the bin tables, nothing more
    idx = lib.msda_index_bytes(16, 5100, 8, 32, 4, 5100, 4)
    assert 0 < idx < (32 << 20)
    assert lib.msda_backward_workspace_bytes(0, 1, 1, 1, 1, 1, 1, 0, 0) == 0


def test_no_cpu_fallback():
    x = make_inputs(N=1, dist="decoder", Lq=3, shapes=[(4, 4)], M=2, D=32)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        MSDeformAttnFunction.apply(x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations,
                                   x.attention_weights, 64)
    m = MSDeformAttn(d_model=64, n_levels=1, n_heads=2, n_points=2)
    q = torch.randn(1, 3, 64)
    ref = torch.rand(1, 3, 1, 2)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        m(q, ref, torch.randn(1, 16, 64), x.spatial_shapes, x.level_start_index)


def test_missing_library_is_loud(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "libmsda_b200.so")
    with pytest.raises(_lib.MSDAError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_module_api_and_state_dict():
    m = MSDeformAttn()
    assert (m.im2col_step, m.d_model, m.n_levels, m.n_heads, m.n_points) == (64, 256, 4, 8, 4)
    assert sorted(m.state_dict()) == sorted(
        f"{n}.{p}" for n in ("sampling_offsets", "attention_weights", "value_proj", "output_proj")
        for p in ("weight", "bias"))
    assert m.sampling_offsets.weight.shape == (256, 256) and m.attention_weights.weight.shape == (128, 256)
    b = m.sampling_offsets.bias.view(8, 4, 4, 2)
    assert torch.allclose(b[0, :, :, 0], torch.tensor([1., 2., 3., 4.]).expand(4, 4))   # head 0 looks along +x
    assert torch.allclose(b[2, :, 3], torch.tensor([0., 4.]).expand(4, 2), atol=1e-6)    # head 2 along +y
    with pytest.raises(ValueError):
        MSDeformAttn(d_model=250, n_heads=8)


def test_synthetic_inputs():
    x = make_inputs(N=2, dist="encoder")
    assert x.value.shape == (2, 5100, 8, 32) and x.sampling_locations.shape == (2, 5100, 8, 4, 4, 2)
    assert x.level_start_index.tolist() == [0, 3840, 4800, 5040]
    assert torch.allclose(x.attention_weights.sum((-1, -2)), torch.ones(2, 5100, 8), atol=1e-5)
    y = make_inputs(N=2, dist="encoder")
    assert torch.equal(x.sampling_locations, y.sampling_locations)
    u = make_inputs(N=1, dist="uniform", seed=1)
    outside = ((u.sampling_locations < 0) | (u.sampling_locations > 1)).any(-1).float().mean()
    assert 0.005 < float(outside) < 0.05
    assert sum(h * w for h, w in scaled_pyramid(20000)) > 15000
    fwd, bwd = algorithmic_bytes(16, 5100, 8, 32, 4, 5100, 4, 4, 4)
    assert (fwd, bwd) == (81600 * 3584, 81600 * 6144)
    fwd, bwd = algorithmic_bytes(16, 5100, 8, 32, 4, 5100, 4, 2, 4)
    assert (fwd, bwd) == (81600 * 2560, 81600 * 4608)


def test_host_pipeline_chunk_plan_and_no_cpu_path():
    from neurips2023_soc_b200 import host_frames
    assert host_frames.chunk_ranges(16, 2) == [(2 * i, 2 * i + 2) for i in range(8)]
    assert host_frames.chunk_ranges(5, 2) == [(0, 2), (2, 4), (4, 5)]
    assert host_frames.chunk_ranges(0, 4) == []
    assert host_frames.chunk_ranges(3, 8) == [(0, 3)]
    with pytest.raises(ValueError):
        host_frames.chunk_ranges(4, 0)
    assert [b - a for a, b in host_frames.ramped_chunk_ranges(16, 4)] == [1, 1, 2, 4, 4, 2, 1, 1]
    for n in range(0, 40):
        for f in (1, 2, 3, 4, 8, 16):
            c = host_frames.ramped_chunk_ranges(n, f)
            assert [a for a, _ in c] == [0] * bool(c) + [b for _, b in c[:-1]]      # contiguous from 0
            assert (c[-1][1] if c else 0) == n and all(0 < b - a <= f for a, b in c)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU path"):
            host_frames.HostFramePipeline("cuda:0")


def _run_bench(*flags, env=None):
    import json
    import os
    import subprocess
    import sys
    e = dict(os.environ, **(env or {}))
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), *flags], capture_output=True, text=True, env=e, timeout=600)
    return r.returncode, [json.loads(ln) for ln in r.stdout.splitlines() if ln.strip()], r.stdout, r.stderr


def test_bench_reference_arm_line_and_bounded_sample():
    """bench.py --impl reference (the CPU formulation on the host cores): one JSON line on stdout with the arm's
    contract keys; a run that would exceed its time bound shrinks a step to fewer frames and says so; ranks other
    than 0 print nothing and exit 0."""
    rc, lines, out, _ = _run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-frames", "2")
    assert rc == 0 and len(lines) == 1 and out.count("\n") == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["metric"] == "msdeformattn_fwd_bwd_queries_per_sec" and d["unit"] == "queries/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["value"] > 0
    # "reference": the reference's own function, staged under baseline/_ref/soc (tools/stage_reference.py); "port": its
    # restatement in oracle/ when nothing was staged
    staged = (ROOT / "baseline" / "_ref" / "soc" / "models" / "ops" / "functions" / "ms_deform_attn_func.py").exists()
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["workload"].startswith("SOC Video-Swin-T deformable encoder") and d["scaling"] == "weak"
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["sample"].startswith("2 of the 16 frames")
    rc, lines, _, _ = _run_bench("--impl", "reference", "--steps", "2", "--warmup", "0", "--ref-frames", "3",
                                 "--ref-budget-s", "0.001")
    assert rc == 0 and lines[0]["cpu_baseline"]["sample"].startswith("1 of the 16 frames")
    rc, lines, out, _ = _run_bench("--impl", "reference", "--steps", "1", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert rc == 0 and out == ""


def test_bench_own_arm_refuses_to_run_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    rc, lines, _, err = _run_bench("--steps", "1")
    assert rc != 0 and not lines and "no CPU fallback" in err


def test_fused_entry_points_refuse_cpu_tensors_and_bad_masks():
    """The fused prologue / encoder-layer entry points of the shim have the reference's error behaviour too: CPU
    tensors are refused before anything is launched (there is no CPU path), and so is a malformed padding mask."""
    N, S, M, D, L, P, Lq = 1, 6, 2, 32, 1, 4, 3
    value = torch.zeros(N, S, M, D)
    shapes = torch.tensor([[2, 3]], dtype=torch.long)
    lsi = torch.tensor([0], dtype=torch.long)
    ref = torch.zeros(N, Lq, L, 2)
    off = torch.zeros(N, Lq, M, L, P, 2)
    logit = torch.zeros(N, Lq, M, L * P)
    go = torch.zeros(N, Lq, M * D)
    with pytest.raises(RuntimeError, match="CUDA tensor|Not implemented on the CPU"):
        msda_ext.ms_deform_attn_forward_fused(value, shapes, lsi, ref, off, logit, 64)
    with pytest.raises(RuntimeError, match="CUDA tensor|Not implemented on the CPU"):
        msda_ext.ms_deform_attn_backward_fused(value, shapes, lsi, off, logit.view(N, Lq, M, L, P), go, 64)
    with pytest.raises(RuntimeError, match="CUDA tensor|Not implemented on the CPU"):
        msda_ext.ms_deform_attn_backward_fused_raw(value, shapes, lsi, ref, off, logit, go, 64, index=torch.zeros(16, dtype=torch.uint8))
    x = torch.zeros(4, 256)
    assert not msda_ext.add_layernorm_supported(x)                      # CPU rows: the caller keeps torch's LayerNorm
    with pytest.raises(RuntimeError, match="CUDA tensor|Not implemented on the CPU"):
        msda_ext.add_layernorm_forward(x, x.clone(), torch.ones(256), torch.zeros(256))
    with pytest.raises(RuntimeError, match="padding_mask"):
        msda_ext._mask_ptr(torch.zeros(N, S, dtype=torch.bool), value)  # a mask must live on the device of value
    assert msda_ext._mask_ptr(None, value) is None
    assert not msda_ext.fused_prologue_supported(value, L, P, 2)        # CPU value: the module keeps the elementwise sequence
