"""Per-kernel device times (CUDA events inside the C ABI, msda_profile_*) of one forward + backward at a given
shape: the A2D encoder shape by default, decoder cross-attention shapes with --lq.  Development tool."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurips2023_soc_b200 import _lib, msda_ext  # noqa: E402
from neurips2023_soc_b200.synthetic import make_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=16)
ap.add_argument("--lq", type=int, default=0, help="queries per frame (decoder distribution); 0: encoder, Lq = S")
ap.add_argument("--dtype", default="bf16mix")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--self-count", action="store_true", help="backward without the forward's index")
ap.add_argument("--flags", type=int, default=0, help="MSDA_FLAG_* bits for every call (A/B switches)")
a = ap.parse_args()
msda_ext.DEFAULT_FLAGS = a.flags
vdt, adt = {"fp32": (torch.float32, torch.float32), "bf16mix": (torch.bfloat16, torch.float32),
            "bf16": (torch.bfloat16, torch.bfloat16)}[a.dtype]
x = (make_inputs(N=a.N, Lq=a.lq, dist="decoder", seed=0) if a.lq else make_inputs(N=a.N, dist="encoder", seed=0))
x = x.to("cuda:0", vdt, adt)
args = (x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights)


def step():
    if a.self_count:
        msda_ext.ms_deform_attn_forward(*args, 64)
        msda_ext.ms_deform_attn_backward(*args, x.grad_output, 64)
    else:
        _, index = msda_ext.ms_deform_attn_forward(*args, 64, want_index=True)
        msda_ext.ms_deform_attn_backward(*args, x.grad_output, 64, index=index)


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
    tot[name] = tot.get(name, 0.0) + ms
print(f"N={a.N} Lq={x.sampling_locations.shape[1]} {a.dtype}: {sum(tot.values()) / a.steps * 1e3:.1f} us per step")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"  {k:40s} {v / a.steps * 1e3:8.1f} us")
