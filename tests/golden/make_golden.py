"""Generate tests/golden/*.npz from the UNMODIFIED reference oracle.

Run in the build container only (needs /root/reference; the GPU box has no
copy of it):

    python tests/golden/make_golden.py

Every case is evaluated by the reference's own
``ms_deform_attn_core_pytorch``
(/root/reference/models/ops/functions/ms_deform_attn_func.py:41-61), imported
from where it lies, in float64 (inputs are float32-representable so the same
file pins both the fp32 and the fp64 paths) plus autograd through it for the
three gradients.  The reference module imports the compiled extension
``MultiScaleDeformableAttention`` at import time (:18); an empty stand-in
module satisfies that import -- nothing from it is called here.

The cases restate the reference's own test shape (models/ops/test.py:21-36,
seed 3) and the edge cases SURVEY.md 8c lists: out-of-range locations, exact
pixel-centre / pixel-edge locations, P=8, L=1, odd channel counts, an
encoder-like pyramid with queries on their own pixels, and a decoder-like
call with few queries.
"""
import importlib.util
import math
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF_FILE = Path("/root/reference/models/ops/functions/ms_deform_attn_func.py")
OUT_DIR = Path(__file__).resolve().parent


def load_reference_oracle():
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    spec = importlib.util.spec_from_file_location("_ref_ms_deform_attn_func", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ms_deform_attn_core_pytorch


def lsi_of(shapes):
    sizes = [h * w for h, w in shapes]
    return [0] + list(np.cumsum(sizes)[:-1])


def normalised_attn(g, N, Lq, M, L, P):
    a = torch.rand(N, Lq, M, L, P, generator=g) + 1e-5
    return a / a.sum(-1, keepdim=True).sum(-2, keepdim=True)


def case_testpy(g):
    # models/ops/test.py:21-36 -- N,M,D = 1,2,2; Lq,L,P = 2,2,2; shapes (6,4),(3,2)
    shapes = [(6, 4), (3, 2)]
    N, M, D, Lq, L, P = 1, 2, 2, 2, 2, 2
    S = sum(h * w for h, w in shapes)
    return dict(shapes=shapes, value=torch.rand(N, S, M, D, generator=g) * 0.01,
                loc=torch.rand(N, Lq, M, L, P, 2, generator=g),
                attn=normalised_attn(g, N, Lq, M, L, P))


def case_out_of_range(g):
    shapes = [(7, 9), (4, 5), (2, 3)]
    N, M, D, Lq, L, P = 2, 3, 8, 13, 3, 4
    S = sum(h * w for h, w in shapes)
    return dict(shapes=shapes, value=torch.randn(N, S, M, D, generator=g),
                loc=torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.6 - 0.3,
                attn=normalised_attn(g, N, Lq, M, L, P))


def case_pixel_lattice(g):
    # locations exactly on pixel centres ((k+.5)/W), on pixel edges (k/W), at 0 and at 1
    shapes = [(4, 6), (2, 3)]
    N, M, D, L, P = 1, 2, 4, 2, 4
    S = sum(h * w for h, w in shapes)
    xs = [0.0, 1.0, 0.5 / 6, 2.5 / 6, 5.5 / 6, 1.0 / 6, 3.0 / 6, 0.5 / 3, 2.5 / 3, 1.0 / 3, -0.5 / 6, 6.5 / 6]
    ys = [0.0, 1.0, 0.5 / 4, 3.5 / 4, 1.0 / 4, 2.0 / 4, 0.5 / 2, 1.5 / 2, 1.0 / 2, -0.5 / 4, 4.5 / 4, 0.25]
    pts = torch.tensor([(x, y) for x in xs for y in ys], dtype=torch.float32)
    Lq = pts.shape[0] // (M * L * P) + 1
    need = Lq * M * L * P
    pts = pts.repeat(math.ceil(need / pts.shape[0]), 1)[:need]
    return dict(shapes=shapes, value=torch.randn(N, S, M, D, generator=g),
                loc=pts.view(N, Lq, M, L, P, 2).clone(),
                attn=normalised_attn(g, N, Lq, M, L, P))


def case_p8_l1(g):
    shapes = [(5, 7)]
    N, M, D, Lq, L, P = 2, 2, 32, 11, 1, 8
    S = 35
    return dict(shapes=shapes, value=torch.randn(N, S, M, D, generator=g),
                loc=torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.2 - 0.1,
                attn=normalised_attn(g, N, Lq, M, L, P))


def case_channels(g, D):
    # channel counts models/ops/test.py:85 walks through (kept to the small ones here)
    shapes = [(6, 4), (3, 2)]
    N, M, Lq, L, P = 1, 2, 3, 2, 2
    S = sum(h * w for h, w in shapes)
    return dict(shapes=shapes, value=torch.rand(N, S, M, D, generator=g) * 0.01,
                loc=torch.rand(N, Lq, M, L, P, 2, generator=g),
                attn=normalised_attn(g, N, Lq, M, L, P))


def case_encoder_like(g):
    # queries = the pyramid's own pixels (deformable_transformer.py:273-285), offsets =
    # the module's compass initialisation (ms_deform_attn.py:65-69) + 1 px noise
    shapes = [(8, 12), (4, 6), (2, 3), (1, 2)]
    N, M, D, L, P = 1, 4, 32, 4, 4
    S = sum(h * w for h, w in shapes)
    ref = []
    for h, w in shapes:
        yy, xx = torch.meshgrid(torch.linspace(0.5, h - 0.5, h) / h, torch.linspace(0.5, w - 0.5, w) / w,
                                indexing="ij")
        ref.append(torch.stack([xx.reshape(-1), yy.reshape(-1)], -1))
    ref = torch.cat(ref, 0)                                                    # (S, 2)
    th = torch.arange(M, dtype=torch.float32) * (2.0 * math.pi / M)
    comp = torch.stack([th.cos(), th.sin()], -1)
    comp = comp / comp.abs().max(-1, keepdim=True)[0]
    off = comp.view(M, 1, 1, 2) * torch.arange(1, P + 1, dtype=torch.float32).view(1, 1, P, 1)
    off = off.expand(M, L, P, 2) + torch.randn(N, S, M, L, P, 2, generator=g)
    norm = torch.tensor([(w, h) for h, w in shapes], dtype=torch.float32).view(1, 1, 1, L, 1, 2)
    loc = ref.view(1, S, 1, 1, 1, 2) + off / norm
    attn = torch.softmax(torch.randn(N, S, M, L * P, generator=g), -1).view(N, S, M, L, P)
    return dict(shapes=shapes, value=torch.randn(N, S, M, D, generator=g), loc=loc.contiguous(), attn=attn)


def case_decoder_like(g):
    shapes = [(8, 12), (4, 6), (2, 3), (1, 2)]
    N, M, D, Lq, L, P = 3, 8, 32, 5, 4, 4
    S = sum(h * w for h, w in shapes)
    ref = torch.sigmoid(torch.randn(N, Lq, 1, 1, 1, 2, generator=g))
    norm = torch.tensor([(w, h) for h, w in shapes], dtype=torch.float32).view(1, 1, 1, L, 1, 2)
    loc = ref + 2.0 * torch.randn(N, Lq, M, L, P, 2, generator=g) / norm
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    value = torch.randn(N, S, M, D, generator=g)
    value[:, -3:] = 0.0                     # padding-masked rows (ms_deform_attn.py:96-97)
    return dict(shapes=shapes, value=value, loc=loc.contiguous(), attn=attn)


def main():
    oracle = load_reference_oracle()
    cases = {
        "testpy_tiny": (3, case_testpy),
        "out_of_range": (11, case_out_of_range),
        "pixel_lattice": (12, case_pixel_lattice),
        "p8_l1": (13, case_p8_l1),
        "channels_30": (14, lambda g: case_channels(g, 30)),
        "channels_71": (15, lambda g: case_channels(g, 71)),
        "channels_64": (16, lambda g: case_channels(g, 64)),
        "encoder_like": (17, case_encoder_like),
        "decoder_like": (18, case_decoder_like),
    }
    for name, (seed, make) in cases.items():
        g = torch.Generator().manual_seed(seed)
        c = make(g)
        shapes = torch.tensor(c["shapes"], dtype=torch.long)
        value, loc, attn = c["value"].float(), c["loc"].float(), c["attn"].float()
        N, Lq = loc.shape[0], loc.shape[1]
        grad_out = torch.randn(N, Lq, value.shape[2] * value.shape[3], generator=g)
        v = value.double().requires_grad_(True)
        lo = loc.double().requires_grad_(True)
        at = attn.double().requires_grad_(True)
        out = oracle(v, shapes, lo, at)
        out.backward(grad_out.double())
        out32 = oracle(value, shapes, loc, attn)
        np.savez_compressed(
            OUT_DIR / f"{name}.npz",
            shapes=shapes.numpy(), lsi=np.asarray(lsi_of(c["shapes"]), dtype=np.int64),
            value=value.numpy(), loc=loc.numpy(), attn=attn.numpy(), grad_out=grad_out.numpy(),
            out=out.detach().numpy(), out_f32=out32.numpy(),
            grad_value=v.grad.numpy(), grad_loc=lo.grad.numpy(), grad_attn=at.grad.numpy())
        print(f"{name}: value {tuple(value.shape)} loc {tuple(loc.shape)} -> out {tuple(out.shape)}")


if __name__ == "__main__":
    main()
