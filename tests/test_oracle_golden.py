"""The CPU oracle against the golden vectors produced by the reference itself.

tests/golden/*.npz hold outputs of the unmodified
/root/reference/models/ops/functions/ms_deform_attn_func.py:41-61 (float64) and
autograd through it (tests/golden/make_golden.py).  Both restatements in
oracle/ must reproduce them before they are trusted as checkers.
"""
import numpy as np
import torch

from oracle import msda_oracle as O


def _close(a, b, tol, keep=None):
    scale = max(1.0, float(np.abs(b).max()))
    err = np.abs(np.asarray(a, dtype=np.float64) - b)
    if keep is not None:
        err = err * keep
    return float(err.max()) <= tol * scale


def kink_free(golden):
    """1 where grad_loc is comparable: not exactly on the accept boundary h_im/w_im == -1
    (cuh:288 rejects the sample; grid_sample autograd returns a one-sided slope)."""
    loc = golden["loc"].astype(np.float64)
    wh = golden["shapes"][:, ::-1].astype(np.float64).reshape(1, 1, 1, -1, 1, 2)
    px = loc * wh - 0.5
    ok = (px != -1.0).all(-1, keepdims=True)
    return np.broadcast_to(ok, loc.shape).astype(np.float64)


def off_lattice(golden, eps=1e-3):
    """1 where the sample sits more than ``eps`` px from an integer pixel coordinate.
    grad_loc is the slope of a piecewise-bilinear surface, discontinuous across pixel
    boundaries, so a reduced-precision evaluation may legitimately pick the other cell."""
    loc = golden["loc"].astype(np.float64)
    wh = golden["shapes"][:, ::-1].astype(np.float64).reshape(1, 1, 1, -1, 1, 2)
    px = loc * wh - 0.5
    ok = (np.abs(px - np.round(px)) > eps).all(-1, keepdims=True)
    return np.broadcast_to(ok, loc.shape).astype(np.float64)


def test_c_oracle_forward_f64(golden):
    out = O.forward_c(golden["value"], golden["shapes"], golden["lsi"], golden["loc"], golden["attn"])
    assert _close(out, golden["out"], 1e-12)


def test_c_oracle_backward_f64(golden):
    gv, gl, ga = O.backward_c(golden["value"], golden["shapes"], golden["lsi"], golden["loc"],
                              golden["attn"], golden["grad_out"])
    assert _close(gv, golden["grad_value"], 1e-12)
    assert _close(gl, golden["grad_loc"], 1e-11, kink_free(golden))
    assert _close(ga, golden["grad_attn"], 1e-12)


def test_c_oracle_f32(golden):
    out = O.forward_c(golden["value"], golden["shapes"], golden["lsi"], golden["loc"], golden["attn"],
                      dtype=np.float32)
    assert _close(out, golden["out"], 1e-5)
    gv, gl, ga = O.backward_c(golden["value"], golden["shapes"], golden["lsi"], golden["loc"],
                              golden["attn"], golden["grad_out"], dtype=np.float32)
    assert _close(gv, golden["grad_value"], 1e-5)
    assert _close(gl, golden["grad_loc"], 2e-5, off_lattice(golden))
    assert _close(ga, golden["grad_attn"], 1e-5)


def test_grid_sample_port(golden):
    t = {k: torch.from_numpy(golden[k]) for k in ("value", "loc", "attn", "grad_out", "shapes")}
    out, gv, gl, ga = O.grid_sample_port_grads(t["value"].double(), t["shapes"], t["loc"].double(),
                                               t["attn"].double(), t["grad_out"].double())
    assert _close(out.numpy(), golden["out"], 1e-12)
    assert _close(gv.numpy(), golden["grad_value"], 1e-12)
    assert _close(gl.numpy(), golden["grad_loc"], 1e-11)
    assert _close(ga.numpy(), golden["grad_attn"], 1e-12)
    out32 = O.grid_sample_port(t["value"], t["shapes"], t["loc"], t["attn"])
    assert _close(out32.numpy(), golden["out_f32"].astype(np.float64), 1e-6)


def test_c_oracle_thread_count_independent():
    g = np.load(__import__("pathlib").Path(__file__).parent / "golden" / "out_of_range.npz")
    args = (g["value"], g["shapes"], g["lsi"], g["loc"], g["attn"], g["grad_out"])
    O.c_oracle_set_threads(1)
    a = O.backward_c(*args, dtype=np.float32)
    O.c_oracle_set_threads(4)
    b = O.backward_c(*args, dtype=np.float32)
    O.c_oracle_set_threads(O.default_threads())
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
