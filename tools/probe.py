"""GPU probe: device-timed forward / backward of this repo's kernels, the atomic A/B arm and the
reference's own CUDA kernels (oracle/_ref, when built) at the A2D shape.  Development tool --
bench.py is the contract; this prints a table and writes gpurun_out/probe.json."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurips2023_soc_b200 import _lib, msda_ext  # noqa: E402
from neurips2023_soc_b200.synthetic import algorithmic_bytes, make_inputs  # noqa: E402


def timed(fn, iters, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=16)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--dists", default="encoder,uniform")
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = "cuda:0"
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ref = None
    if not args.no_ref:
        from oracle import build_ref
        ref = build_ref.load()
    rows = []
    for dist in args.dists.split(","):
        base = make_inputs(N=args.N, dist=dist, seed=0)
        for tag, vdt, adt in (("fp32", torch.float32, torch.float32), ("bf16mix", torch.bfloat16, torch.float32),
                              ("bf16", torch.bfloat16, torch.bfloat16)):
            if args.only and tag not in args.only.split(","):
                continue
            x = base.to(dev, vdt, adt)
            a = (x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights)
            N, S, M, D = x.value.shape
            Lq, L, P = x.sampling_locations.shape[1], 4, x.sampling_locations.shape[4]
            fb, bb = algorithmic_bytes(N, S, M, D, L, Lq, P, x.value.element_size(), x.sampling_locations.element_size())
            f_med, f_min = timed(lambda: msda_ext.ms_deform_attn_forward(*a, 64), args.iters, flush)
            fi_med, _ = timed(lambda: msda_ext.ms_deform_attn_forward(*a, 64, want_index=True), args.iters, flush)

            def pair(flags=None):   # the backward consumes the index, so forward and backward are timed as a pair
                _, index = msda_ext.ms_deform_attn_forward(*a, 64, want_index=True)
                msda_ext.ms_deform_attn_backward(*a, x.grad_output, 64, index=index, flags=flags)
            p_med, _ = timed(pair, args.iters, flush)
            b_med = p_med - fi_med
            bs_med, _ = timed(lambda: msda_ext.ms_deform_attn_backward(*a, x.grad_output, 64), args.iters, flush)
            row = dict(dist=dist, dtype=tag, N=N, fwd_us=f_med, fwd_indexed_us=fi_med, bwd_us=b_med, bwd_selfcount_us=bs_med,
                       fwd_GBs=fb / f_med / 1e3, bwd_GBs=bb / b_med / 1e3,
                       fwdbwd_us=p_med, fwdbwd_Mq_s=N * Lq / p_med, launches_bwd=msda_ext.last_launch_count())
            if tag == "fp32":
                at_med, _ = timed(lambda: msda_ext.ms_deform_attn_backward(*a, x.grad_output, 64,
                                                                           flags=_lib.FLAG_ATOMIC_GRAD_VALUE), args.iters, flush)
                lin_med, _ = timed(lambda: msda_ext.ms_deform_attn_forward(*a, 64, flags=_lib.FLAG_PYRAMID_TILES), args.iters, flush)
                gen_med, _ = timed(lambda: msda_ext.ms_deform_attn_forward(*a, 64, flags=_lib.FLAG_GENERIC), max(3, args.iters // 4), flush)
                row.update(bwd_atomic_us=at_med, fwd_pyramid_tiles_us=lin_med, fwd_generic_us=gen_med)
                if ref is not None:
                    rf, _ = timed(lambda: ref.ms_deform_attn_forward(*a, 64), max(3, args.iters // 2), flush)
                    rb, _ = timed(lambda: ref.ms_deform_attn_backward(*a, x.grad_output, 64), max(3, args.iters // 2), flush)
                    row.update(ref_cuda_fwd_us=rf, ref_cuda_bwd_us=rb)
            rows.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe.json", "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
