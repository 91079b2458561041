"""Run a few forward+backward calls of the op at the A2D shape (for ncu)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurips2023_soc_b200 import msda_ext  # noqa: E402
from neurips2023_soc_b200.synthetic import make_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=16)
ap.add_argument("--dtype", default="fp32")
ap.add_argument("--dist", default="encoder")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--fwd-only", action="store_true")
a = ap.parse_args()
vdt, adt = {"fp32": (torch.float32, torch.float32), "bf16mix": (torch.bfloat16, torch.float32),
            "bf16": (torch.bfloat16, torch.bfloat16)}[a.dtype]
x = make_inputs(N=a.N, dist=a.dist, seed=0).to("cuda:0", vdt, adt)
args = (x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights)
for _ in range(a.steps):
    if a.fwd_only:
        msda_ext.ms_deform_attn_forward(*args, 64)
    else:
        _, index = msda_ext.ms_deform_attn_forward(*args, 64, want_index=True)
        msda_ext.ms_deform_attn_backward(*args, x.grad_output, 64, index=index)
torch.cuda.synchronize()
