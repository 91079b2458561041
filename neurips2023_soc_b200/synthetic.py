"""Seeded synthetic inputs for the MSDeformAttn hot path.

Shapes follow the reference configs: ``d_model 256, nheads 8,
num_feature_levels 4, enc/dec_n_points 4``
(/root/reference/configs/a2d_sentences.yaml:36-63) and the 384x640 Video-Swin
pyramid 48x80 / 24x40 / 12x20 / 6x10.

Two location distributions, both reported by the bench:

* ``"encoder"`` -- every query is a pixel of the pyramid; its reference point is
  that pixel's centre mapped to all levels
  (``DeformableTransformerEncoder.get_reference_points``,
  /root/reference/models/deformable_transformer.py:273-285, valid_ratio 1) and
  the offsets are the module's compass initialisation
  (/root/reference/models/ops/modules/ms_deform_attn.py:63-71: direction of
  head ``m`` times ``p+1`` pixels) plus N(0, 1 px) noise.
* ``"uniform"``  -- ``rand in [0,1)^2`` as /root/reference/models/ops/test.py:34,
  with 2 % of the samples pushed outside [0,1] to exercise zero padding.
* ``"decoder"``  -- few queries per frame with ``sigmoid(N(0,1))`` reference
  points and 2 px offset noise (decoder cross-attention,
  /root/reference/models/deformable_transformer.py:330-347).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import torch

A2D_PYRAMID: Tuple[Tuple[int, int], ...] = ((48, 80), (24, 40), (12, 20), (6, 10))


@dataclass
class MSDAInputs:
    value: torch.Tensor            # (N, S, M, D)
    spatial_shapes: torch.Tensor   # (L, 2) int64 (H, W)
    level_start_index: torch.Tensor  # (L,) int64
    sampling_locations: torch.Tensor  # (N, Lq, M, L, P, 2)
    attention_weights: torch.Tensor   # (N, Lq, M, L, P)
    grad_output: torch.Tensor      # (N, Lq, M*D)

    def to(self, device, value_dtype=None, aux_dtype=None, non_blocking=False):
        vd = value_dtype or self.value.dtype
        ad = aux_dtype or self.sampling_locations.dtype
        return MSDAInputs(
            self.value.to(device=device, dtype=vd, non_blocking=non_blocking),
            self.spatial_shapes.to(device=device, non_blocking=non_blocking),
            self.level_start_index.to(device=device, non_blocking=non_blocking),
            self.sampling_locations.to(device=device, dtype=ad, non_blocking=non_blocking),
            self.attention_weights.to(device=device, dtype=ad, non_blocking=non_blocking),
            self.grad_output.to(device=device, dtype=vd, non_blocking=non_blocking))

    @property
    def num_queries(self) -> int:
        return self.sampling_locations.shape[0] * self.sampling_locations.shape[1]


def level_start_index(shapes: Sequence[Tuple[int, int]]) -> List[int]:
    out, acc = [], 0
    for h, w in shapes:
        out.append(acc)
        acc += h * w
    return out


def scaled_pyramid(tokens: int, levels: int = 4) -> List[Tuple[int, int]]:
    """A 3:5 pyramid with about ``tokens`` tokens in total (config-5 sweep)."""
    # S = sum_l (3k/2^l)(5k/2^l) ~= 15 k^2 * 4/3  ->  k = sqrt(tokens / 20)
    k = max(1 << (levels - 1), int(round(math.sqrt(tokens / 20.0) / (1 << (levels - 1)))) * (1 << (levels - 1)))
    return [(max(1, 3 * k >> l), max(1, 5 * k >> l)) for l in range(levels)]


def pyramid_reference_points(shapes: Sequence[Tuple[int, int]]) -> torch.Tensor:
    """(S, 2) pixel centres (x, y) in [0,1], level-major then row-major."""
    pts = []
    for h, w in shapes:
        ys = (torch.arange(h, dtype=torch.float32) + 0.5) / h
        xs = (torch.arange(w, dtype=torch.float32) + 0.5) / w
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack([xx.reshape(-1), yy.reshape(-1)], -1))
    return torch.cat(pts, 0)


def compass_offsets(M: int, L: int, P: int) -> torch.Tensor:
    """(M, L, P, 2) pixel offsets of ``MSDeformAttn._reset_parameters``."""
    th = torch.arange(M, dtype=torch.float32) * (2.0 * math.pi / M)
    d = torch.stack([th.cos(), th.sin()], -1)
    d = d / d.abs().max(-1, keepdim=True)[0]
    steps = torch.arange(1, P + 1, dtype=torch.float32).view(1, 1, P, 1)
    return (d.view(M, 1, 1, 2) * steps).expand(M, L, P, 2).contiguous()


def make_inputs(N: int = 1, shapes: Sequence[Tuple[int, int]] = A2D_PYRAMID, M: int = 8, D: int = 32,
                P: int = 4, Lq: int | None = None, dist: str = "encoder", seed: int = 0,
                value_scale: float = 1.0) -> MSDAInputs:
    """fp32 CPU tensors; use ``.to(device, value_dtype, aux_dtype)`` afterwards."""
    g = torch.Generator().manual_seed(seed)
    shapes = [(int(h), int(w)) for h, w in shapes]
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    norm = torch.tensor([(w, h) for h, w in shapes], dtype=torch.float32).view(1, 1, 1, L, 1, 2)
    if dist == "encoder":
        Lq = S
        ref = pyramid_reference_points(shapes).view(1, S, 1, 1, 1, 2)
        off = compass_offsets(M, L, P).view(1, 1, M, L, P, 2) + torch.randn(N, Lq, M, L, P, 2, generator=g)
        loc = ref + off / norm
    elif dist == "uniform":
        Lq = S if Lq is None else Lq
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g)
        push = torch.rand(N, Lq, M, L, P, 1, generator=g) < 0.02
        loc = torch.where(push, loc * 1.5 - 0.25 + torch.sign(loc - 0.5), loc)
    elif dist == "decoder":
        Lq = 20 if Lq is None else Lq
        ref = torch.sigmoid(torch.randn(N, Lq, 1, 1, 1, 2, generator=g))
        loc = ref + 2.0 * torch.randn(N, Lq, M, L, P, 2, generator=g) / norm
    else:
        raise ValueError(f"unknown location distribution {dist!r}")
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    value = torch.randn(N, S, M, D, generator=g) * value_scale
    grad_out = torch.randn(N, Lq, M * D, generator=g)
    return MSDAInputs(value.contiguous(), torch.tensor(shapes, dtype=torch.long),
                      torch.tensor(level_start_index(shapes), dtype=torch.long),
                      loc.contiguous(), attn.contiguous(), grad_out.contiguous())


def algorithmic_bytes(N: int, S: int, M: int, D: int, L: int, Lq: int, P: int,
                      value_bytes: int, aux_bytes: int) -> Tuple[int, int]:
    """Compulsory HBM bytes of one forward / one backward call."""
    C = M * D
    samples = N * Lq * M * L * P
    fwd = value_bytes * N * S * C + aux_bytes * samples * 3 + value_bytes * N * Lq * C
    bwd = (value_bytes * N * Lq * C + value_bytes * N * S * C + aux_bytes * samples * 3
           + value_bytes * N * S * C + aux_bytes * samples * 3)
    return fwd, bwd
