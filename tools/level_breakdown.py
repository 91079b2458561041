"""Per-level split of the kernel times at the A2D encoder shape.  Development tool.

The bench step (N = 16 frames, 5100 queries per frame, 4 levels x 4 points) is cut into four single-level
problems that do exactly the work the full call does in that level: the level's value rows, the same queries,
the level's 4 sampling points per (query, head) with their weights, the same grad_output.  Per-kernel device
times come from the C ABI's own CUDA events (msda_profile_*).  The forward / sample-gradient kernels pay their
per-item overhead four times this way, so their rows are upper bounds; the index kernels (sort, walk) work level
by level in the full call too, so their rows add up to the full call's."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurips2023_soc_b200 import _lib, msda_ext  # noqa: E402
from neurips2023_soc_b200.synthetic import make_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=16)
ap.add_argument("--dtype", default="bf16mix")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--dist", default="encoder")
ap.add_argument("--flags", type=int, default=0, help="msda_*_ex flags (32: first-generation sort + walk)")
a = ap.parse_args()
vdt, adt = {"fp32": (torch.float32, torch.float32), "bf16mix": (torch.bfloat16, torch.float32)}[a.dtype]
full = make_inputs(N=a.N, dist=a.dist, seed=0)
dev = "cuda:0"


def times(value, shapes, lsi, loc, attn, gout):
    args = (value, shapes, lsi, loc, attn)

    def step():
        _, index = msda_ext.ms_deform_attn_forward(*args, 64, want_index=True, flags=a.flags)
        msda_ext.ms_deform_attn_backward(*args, gout, 64, index=index, flags=a.flags)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
    recs = _lib.profile_read()
    _lib.profile_enable(False)
    tot = {}
    for name, ms in recs:
        tot[name] = tot.get(name, 0.0) + ms * 1e3 / a.steps
    return {k: round(v, 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])}


x = full.to(dev, vdt, adt)
rows = [{"level": "all", "shape": [list(map(int, s)) for s in full.spatial_shapes.tolist()],
         "us": times(x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights, x.grad_output)}]
for l, (h, w) in enumerate(full.spatial_shapes.tolist()):
    s0 = int(full.level_start_index[l])
    value = x.value[:, s0:s0 + h * w].contiguous()
    loc = x.sampling_locations[:, :, :, l:l + 1].contiguous()
    attn = x.attention_weights[:, :, :, l:l + 1].contiguous()
    shapes = torch.tensor([[h, w]], dtype=torch.long, device=dev)
    lsi = torch.zeros(1, dtype=torch.long, device=dev)
    rows.append({"level": l, "shape": [h, w], "entries_per_bin": round(loc.shape[1] * loc.shape[4] / ((h + 1) * (w + 1)), 1),
                 "us": times(value, shapes, lsi, loc, attn, x.grad_output)})
for r in rows:
    print(json.dumps(r))
