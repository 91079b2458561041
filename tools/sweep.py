"""BASELINE.json configs[4]: MSDeformAttn sweep -- tokens/frame 1k..20k, points 4 and 8, fp32 and bf16,
forward and backward, against the reference's own CUDA kernels rebuilt for sm_100a (oracle/_ref, fp32)
and the reference's CPU formulation (one frame, all host threads).  Development / evidence tool:
writes gpurun_out/sweep.jsonl and prints a markdown table; bench.py is the contract.

Every case is also checked against the reference CUDA op (fp32, max-abs error) so that the table is a
parity statement as well as a timing one."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurips2023_soc_b200 import msda_ext  # noqa: E402
from neurips2023_soc_b200.synthetic import algorithmic_bytes, make_inputs  # noqa: E402
from tools.probe import timed  # noqa: E402


def cpu_times(x, threads):
    from oracle import msda_oracle
    torch.set_num_threads(threads)
    best_f, best_fb = float("inf"), float("inf")
    for _ in range(2):
        v = x.value.clone().requires_grad_(True)
        lo = x.sampling_locations.clone().requires_grad_(True)
        at = x.attention_weights.clone().requires_grad_(True)
        t0 = time.perf_counter()
        out = msda_oracle.grid_sample_port(v, x.spatial_shapes, lo, at)
        t1 = time.perf_counter()
        out.backward(x.grad_output)
        t2 = time.perf_counter()
        best_f, best_fb = min(best_f, t1 - t0), min(best_fb, t2 - t0)
    return best_f, best_fb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=16)
    ap.add_argument("--scales", default="7,10,16,22,32",
                    help="pyramid scale k: level l is ceil(3k/2^l) x ceil(5k/2^l); 7 -> 1002 tokens, 10 -> 2007, "
                         "16 -> 5100 (the A2D pyramid), 22 -> 9677, 32 -> 20400")
    ap.add_argument("--points", default="4,8")
    ap.add_argument("--iters", type=int, default=7)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--decoder", action="store_true", help="also time decoder cross-attention shapes (Lq = 5 / 20 / 300)")
    args = ap.parse_args()
    dev = "cuda:0"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ref = None
    if not args.no_ref:
        try:
            from oracle import build_ref
            ref = build_ref.load()
        except Exception as e:  # evidence tool: the reference arm is optional
            print(f"# reference CUDA op unavailable: {e}")
    threads = os.cpu_count() or 1
    os.makedirs("gpurun_out", exist_ok=True)
    rows = []
    for k in map(int, args.scales.split(",")):
        shapes = [(-(-3 * k // (1 << l)), -(-5 * k // (1 << l))) for l in range(4)]
        tokens = sum(h * w for h, w in shapes)
        for P in map(int, args.points.split(",")):
            base = make_inputs(N=args.N, shapes=shapes, P=P, dist="encoder", seed=tokens + P)
            S = base.value.shape[1]
            cpu_f = cpu_fb = None
            if not args.no_cpu:
                one = make_inputs(N=1, shapes=shapes, P=P, dist="encoder", seed=tokens + P)
                cpu_f, cpu_fb = cpu_times(one, threads)
            ref_res = None
            for tag, vdt, adt in (("fp32", torch.float32, torch.float32), ("bf16", torch.bfloat16, torch.float32)):
                x = base.to(dev, vdt, adt)
                a = (x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights)
                fb, bb = algorithmic_bytes(args.N, S, 8, 32, len(shapes), S, P, x.value.element_size(), 4)
                f_med, _ = timed(lambda: msda_ext.ms_deform_attn_forward(*a, 64), args.iters, flush)
                fi_med, _ = timed(lambda: msda_ext.ms_deform_attn_forward(*a, 64, want_index=True), args.iters, flush)

                def pair():
                    _, index = msda_ext.ms_deform_attn_forward(*a, 64, want_index=True)
                    return msda_ext.ms_deform_attn_backward(*a, x.grad_output, 64, index=index)
                p_med, _ = timed(pair, args.iters, flush)
                row = dict(tokens=S, shapes=shapes, P=P, dtype=tag, N=args.N, fwd_us=f_med, bwd_us=p_med - fi_med,
                           fwdbwd_us=p_med, fwdbwd_Mq_s=args.N * S / p_med,
                           hbm_frac_fwdbwd=(fb + bb) / (p_med * 1e-6) / 6553e9)
                if tag == "fp32" and ref is not None:
                    rf, _ = timed(lambda: ref.ms_deform_attn_forward(*a, 64), max(3, args.iters // 2), flush)
                    rb, _ = timed(lambda: ref.ms_deform_attn_backward(*a, x.grad_output, 64), max(3, args.iters // 2), flush)
                    out = msda_ext.ms_deform_attn_forward(*a, 64)
                    grads = pair()
                    r_out = ref.ms_deform_attn_forward(*a, 64)
                    r_grads = ref.ms_deform_attn_backward(*a, x.grad_output, 64)
                    errs = [float((g - r).abs().max() / max(1.0, float(r.abs().max())))
                            for g, r in zip([out] + list(grads), [r_out] + list(r_grads))]
                    row.update(ref_cuda_fwd_us=rf, ref_cuda_bwd_us=rb, speedup_fwd=rf / f_med,
                               speedup_bwd=rb / (p_med - fi_med), max_err_vs_ref_cuda=max(errs))
                    ref_res = (rf, rb)
                if cpu_f is not None:
                    row.update(cpu_fwd_ms_per_frame=cpu_f * 1e3, cpu_fwdbwd_ms_per_frame=cpu_fb * 1e3, cpu_threads=threads,
                               speedup_vs_cpu_fwdbwd=(cpu_fb * args.N) / (p_med * 1e-6))
                rows.append(row)
                print(json.dumps(row), flush=True)
                del x, a
            del base
            torch.cuda.empty_cache()
    # decoder cross-attention shapes (BASELINE.json configs[1] and [3]): few queries per frame over the A2D pyramid
    if args.decoder:
        for N, Lq in ((16, 20), (2, 20), (36, 5), (36, 20), (16, 300)):
            base = make_inputs(N=N, Lq=Lq, dist="decoder", seed=Lq)
            S = base.value.shape[1]
            for tag, vdt, adt in (("fp32", torch.float32, torch.float32), ("bf16", torch.bfloat16, torch.float32)):
                x = base.to(dev, vdt, adt)
                a = (x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights)
                f_med, _ = timed(lambda: msda_ext.ms_deform_attn_forward(*a, 64), args.iters, flush)
                fi_med, _ = timed(lambda: msda_ext.ms_deform_attn_forward(*a, 64, want_index=True), args.iters, flush)

                def pair():
                    _, index = msda_ext.ms_deform_attn_forward(*a, 64, want_index=True)
                    return msda_ext.ms_deform_attn_backward(*a, x.grad_output, 64, index=index)
                p_med, _ = timed(pair, args.iters, flush)
                row = dict(tokens=S, P=4, dtype=tag, N=N, Lq=Lq, kind="decoder", fwd_us=f_med, bwd_us=p_med - fi_med,
                           fwdbwd_us=p_med, fwdbwd_Mq_s=N * Lq / p_med)
                if tag == "fp32" and ref is not None:
                    rf, _ = timed(lambda: ref.ms_deform_attn_forward(*a, 64), args.iters, flush)
                    rb, _ = timed(lambda: ref.ms_deform_attn_backward(*a, x.grad_output, 64), args.iters, flush)
                    out = msda_ext.ms_deform_attn_forward(*a, 64)
                    grads = pair()
                    r_out = ref.ms_deform_attn_forward(*a, 64)
                    r_grads = ref.ms_deform_attn_backward(*a, x.grad_output, 64)
                    errs = [float((g - r).abs().max() / max(1.0, float(r.abs().max())))
                            for g, r in zip([out] + list(grads), [r_out] + list(r_grads))]
                    row.update(ref_cuda_fwd_us=rf, ref_cuda_bwd_us=rb, speedup_fwd=rf / f_med,
                               speedup_bwd=rb / (p_med - fi_med), max_err_vs_ref_cuda=max(errs))
                rows.append(row)
                print(json.dumps(row), flush=True)
    with open("gpurun_out/sweep.jsonl", "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
    print("\n| tokens | P | dtype | fwd us | bwd us | M queries/s | ref CUDA fwd us | ref CUDA bwd us | x fwd | x bwd | "
          "max err vs ref CUDA | CPU fwd+bwd ms/frame | x CPU |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        g = lambda k, f="%.0f": (f % r[k]) if k in r and r[k] is not None else "-"   # noqa: E731
        if r.get("kind") == "decoder":
            r = dict(r, tokens=f"{r['tokens']} (N={r['N']}, Lq={r['Lq']})")
        print(f"| {r['tokens']} | {r['P']} | {r['dtype']} | {g('fwd_us')} | {g('bwd_us')} | {g('fwdbwd_Mq_s', '%.1f')} | "
              f"{g('ref_cuda_fwd_us')} | {g('ref_cuda_bwd_us')} | {g('speedup_fwd', '%.1f')} | {g('speedup_bwd', '%.1f')} | "
              f"{g('max_err_vs_ref_cuda', '%.1e')} | {g('cpu_fwdbwd_ms_per_frame', '%.0f')} | {g('speedup_vs_cpu_fwdbwd', '%.0f')} |")


if __name__ == "__main__":
    main()
