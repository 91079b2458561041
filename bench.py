#!/usr/bin/env python
"""bench.py -- MSDeformAttn forward + backward at the A2D-Sentences Video-Swin-T encoder shape.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 8-frame clips, batch 2 => 16 frames per GPU, 5100 tokens per
frame (48x80 / 24x40 / 12x20 / 6x10), d_model 256 = 8 heads x 32, 4 levels x 4 points, queries =
tokens (encoder self-attention).  bf16 values / outputs / value gradients, fp32 sampling locations
and attention weights (what autocast produces), fp32 accumulation.  One STEP = one forward + one
backward of the op over the rank's 16 frames.  Frames are independent, so ranks hold disjoint
frames and there is no data-path collective (weak scaling); `value` = queries of all ranks / time.

Printed JSON line (rank 0): metric/value/unit..., plus
  roofline     -- the dominant kernel (longest device time per step, measured live with CUDA events
                  through msda_profile_*), its algorithmic bytes per launch over its
                  average duration, against MEASURED_PEAKS.json's copy bandwidth
  cpu_baseline -- the reference's CPU formulation (grid_sample; oracle/msda_oracle.py) timed on this
                  box's host cores on one frame of the same workload
  e2e          -- the same step through the host entry point (host_frames.HostFramePipeline) with inputs in
                  pinned HOST memory: H2D of value/locations/weights/grad_output and D2H of output + the
                  three gradients are inside the timed region, pipelined over chunks of frames; the
                  one-stream time (MSDeformAttnFunction between blocking-order copies) is beside it
  --impl reference times that CPU formulation alone (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

FRAMES_PER_GPU = 16
METRIC = "msdeformattn_fwd_bwd_queries_per_sec"
UNIT = "queries/s"
WORKLOAD = ("SOC Video-Swin-T deformable encoder, A2D-Sentences shape: 16 frames/GPU (8-frame clips x batch 2), "
            "5100 tokens/frame (48x80,24x40,12x20,6x10), 8 heads x 32, 4 levels x 4 points, Lq = S")


# stdout carries the result line and nothing else: libraries that print there (NCCL's version banner under
# NCCL_DEBUG=VERSION/WARN, for one) are sent to stderr for the duration of the run
_RESULT_FD = None


def claim_stdout():
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj) -> None:
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, line)


def peaks():
    try:
        with open(ROOT / "MEASURED_PEAKS.json") as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(len(r) > col and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
def cpu_formulation(threads: int):
    """(callable, kind): the reference's CPU path (ms_deform_attn_core_pytorch,
    /root/reference/models/ops/functions/ms_deform_attn_func.py:41-61) as restated in oracle/ -- the
    reference tree does not exist on the GPU box, and the restatement is pinned to it by
    tests/test_oracle_golden.py.  All host threads."""
    torch.set_num_threads(threads)
    from oracle import msda_oracle
    return msda_oracle.grid_sample_port, "port"


def time_cpu(fn, x, repeats: int, warm: bool = True):
    """fwd + autograd bwd of one frame on the host; seconds per pass (best of `repeats` after a warm-up,
    or the single pass when warm is False)."""
    best = float("inf")
    for i in range(repeats + 1 if warm else 1):
        v = x.value.clone().requires_grad_(True)
        lo = x.sampling_locations.clone().requires_grad_(True)
        at = x.attention_weights.clone().requires_grad_(True)
        t0 = time.perf_counter()
        out = fn(v, x.spatial_shapes, lo, at)
        out.backward(x.grad_output)
        dt = time.perf_counter() - t0
        if i or not warm:          # with warm, the first pass is the warm-up
            best = min(best, dt)
    return best


def run_reference(args):
    from neurips2023_soc_b200.synthetic import make_inputs
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    fn, kind = cpu_formulation(threads)
    x = make_inputs(N=args.ref_frames, dist="encoder", seed=0)
    passes = args.warmup + max(1, args.steps)
    # bounded sample: one untimed pass over the whole step sizes the run; when `passes` of them would not end
    # within --ref-budget-s, a step becomes the first `frames` frames of the same workload (frames are independent,
    # queries/s is per frame) -- never fewer than one frame
    t_full = time_cpu(fn, x, 0, warm=False)
    frames = args.ref_frames
    if passes * t_full > args.ref_budget_s:
        frames = max(1, min(args.ref_frames, int(args.ref_frames * args.ref_budget_s / (passes * t_full))))
        x = make_inputs(N=frames, dist="encoder", seed=0)
    ts = [time_cpu(fn, x, 0, warm=False) for _ in range(passes)][args.warmup:]
    sec = sum(ts) / len(ts)
    qps = x.num_queries / sec
    sample = (f"{frames} of the 16 frames per step ({x.num_queries} queries), fp32, forward + autograd backward "
              f"through F.grid_sample on {threads} host threads")
    emit({
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from neurips2023_soc_b200 import MSDeformAttnFunction, _lib, msda_ext
    from neurips2023_soc_b200.synthetic import algorithmic_bytes, make_inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the op (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    vdt, adt = torch.bfloat16, torch.float32
    host = make_inputs(N=FRAMES_PER_GPU, dist="encoder", seed=rank)      # each rank owns different frames
    x = host.to(dev, vdt, adt)
    args_t = (x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights)
    N, S, M, D = x.value.shape
    Lq, L, P = x.sampling_locations.shape[1], x.sampling_locations.shape[3], x.sampling_locations.shape[4]
    queries = N * Lq

    def step():
        # what MSDeformAttnFunction does: the forward hands the sub-bin offsets of the inverse index on
        out, index = msda_ext.ms_deform_attn_forward(*args_t, 64, want_index=True)
        lf = msda_ext.last_launch_count()
        grads = msda_ext.ms_deform_attn_backward(*args_t, x.grad_output, 64, index=index)
        return out, grads, lf + msda_ext.last_launch_count()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        _, _, launches_per_step = step()
    sync_all()

    # ---- timed region: K steps, device time on the launching (current) stream ----
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with ClockSampler(local) as clk:
        sync_all()
        ev[0].record()
        for _ in range(args.steps):
            step()
        ev[1].record()
        sync_all()
    ms_total = ev[0].elapsed_time(ev[1])
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * queries / (ms_step * 1e-3)

    # ---- per-kernel device times over the same K steps (instrumented pass) ----
    _lib.profile_enable(True)
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    recs = _lib.profile_read()
    _lib.profile_enable(False)
    per_kernel = {}
    for name, ms in recs:
        per_kernel.setdefault(name, []).append(ms)
    kern_avg = {k: sum(v) / len(v) for k, v in per_kernel.items()}
    kern_per_step = {k: sum(v) / args.steps for k, v in per_kernel.items()}
    dominant = max(kern_per_step, key=kern_per_step.get)
    vb, ab = x.value.element_size(), x.sampling_locations.element_size()
    samples = N * Lq * M * L * P
    C = M * D
    fwd_bytes, bwd_bytes = algorithmic_bytes(N, S, M, D, L, Lq, P, vb, ab)
    # algorithmic bytes of each kernel's own job, per launch
    alg = {
        "msda_fwd_tile_kernel": fwd_bytes,
        "msda_bwd_sample_tile_kernel": vb * N * S * C + vb * N * Lq * C + ab * samples * 3 + ab * samples * 3 + 16 * samples,
        "msda_grad_value_walk_kernel": vb * N * Lq * C + 16 * samples + vb * N * S * C,
        "msda_bin_rank_sort_kernel": 2 * 16 * samples,
    }
    peak, peak_src = peaks()
    dom_bytes = alg.get(dominant, bwd_bytes)
    achieved = dom_bytes / (kern_avg[dominant] * 1e-3) / 1e9
    device_ms = sum(kern_per_step.values())
    roofline = {
        "bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_ms": kern_avg[dominant],
        "kernel_ms_per_step": {k: round(v, 4) for k, v in sorted(kern_per_step.items(), key=lambda kv: -kv[1])},
        "step": {"algorithmic_bytes": fwd_bytes + bwd_bytes, "device_ms": device_ms,
                 "achieved": (fwd_bytes + bwd_bytes) / (device_ms * 1e-3) / 1e9,
                 "frac": (fwd_bytes + bwd_bytes) / (device_ms * 1e-3) / 1e9 / peak},
    }
    traffic_file = ROOT / "profiles" / "dominant_kernel_traffic.json"
    if traffic_file.exists():
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(dominant)
        except Exception:
            pass

    # ---- end to end through the public autograd API with host buffers ----
    pin = {k: getattr(host, k).to(vdt if k in ("value", "grad_output") else adt).pin_memory()
           for k in ("value", "sampling_locations", "attention_weights", "grad_output")}
    res_host = {
        "out": torch.empty((N, Lq, C), dtype=vdt).pin_memory(),
        "gv": torch.empty_like(pin["value"]).pin_memory(),
        "gl": torch.empty_like(pin["sampling_locations"]).pin_memory(),
        "ga": torch.empty_like(pin["attention_weights"]).pin_memory(),
    }
    h2d = sum(t_.numel() * t_.element_size() for t_ in pin.values())
    d2h = sum(t_.numel() * t_.element_size() for t_ in res_host.values())

    def e2e_unpipelined():      # one stream: H2D, the autograd function, D2H back to back
        v = pin["value"].to(dev, non_blocking=True).requires_grad_(True)
        lo = pin["sampling_locations"].to(dev, non_blocking=True).requires_grad_(True)
        at = pin["attention_weights"].to(dev, non_blocking=True).requires_grad_(True)
        go = pin["grad_output"].to(dev, non_blocking=True)
        out = MSDeformAttnFunction.apply(v, x.spatial_shapes, x.level_start_index, lo, at, 64)
        out.backward(go)
        res_host["out"].copy_(out.detach(), non_blocking=True)
        res_host["gv"].copy_(v.grad, non_blocking=True)
        res_host["gl"].copy_(lo.grad, non_blocking=True)
        res_host["ga"].copy_(at.grad, non_blocking=True)

    # the host entry point: frames streamed in chunks, H2D / kernels / D2H on three streams
    from neurips2023_soc_b200.host_frames import HostFramePipeline
    pipe = HostFramePipeline(dev, frames_per_chunk=args.e2e_frames_per_chunk)
    results = (res_host["out"], res_host["gv"], res_host["gl"], res_host["ga"])

    def e2e_step():
        pipe.forward_backward(pin["value"], host.spatial_shapes, host.level_start_index, pin["sampling_locations"],
                              pin["attention_weights"], pin["grad_output"], results=results)

    def time_e2e(fn, steps):
        for _ in range(2):
            fn()
        sync_all()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        evs[0].record()
        for _ in range(steps):
            fn()
        evs[1].record()
        sync_all()
        tt = torch.tensor([evs[0].elapsed_time(evs[1])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()) / steps

    e2e_steps = max(2, min(args.steps, 20))
    e2e_ms = time_e2e(e2e_step, e2e_steps)
    e2e_serial_ms = time_e2e(e2e_unpipelined, max(2, min(args.steps, 5)))
    e2e = {"value": world * queries / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps,
           "api": f"HostFramePipeline.forward_backward: pinned host buffers, chunks of up to {args.e2e_frames_per_chunk} "
                  f"frames (short first and last chunks), H2D / kernels / D2H pipelined on three streams ({pipe.launches} kernel launches per step)",
           "one_stream_ms_per_step": e2e_serial_ms,
           "pcie_GBs_each_way": max(h2d, d2h) / (e2e_ms * 1e-3) / 1e9}

    # ---- CPU baseline beside it (rank 0, N = 1 only): one frame of the same workload ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        fn, kind = cpu_formulation(threads)
        one = make_inputs(N=args.ref_frames, dist="encoder", seed=0)
        sec = time_cpu(fn, one, args.cpu_repeats)
        cpu_baseline = {"value": one.num_queries / sec, "unit": UNIT, "cores": threads, "kind": kind,
                        "sample": f"{args.ref_frames} of the 16 frames ({one.num_queries} queries), fp32, forward + autograd "
                                  f"backward through F.grid_sample, best of {args.cpu_repeats} after 1 warm-up",
                        "ms": sec * 1e3}

    if rank == 0:
        emit({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": N, "queries_per_step_per_gpu": queries,
                       "value_dtype": "bf16", "location_weight_dtype": "f32", "accumulate": "f32",
                       "locations": "encoder-realistic (SURVEY.md 8d distribution A)",
                       "l2": "per-step working set ~0.7 GB (inputs 0.25 GB + outputs/gradients 0.25 GB + index 0.25 GB) "
                             "exceeds the 126 MB L2; no explicit flush",
                       "parallelism": f"frames sharded over {world} rank(s), no data-path collective"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
            "clocks": clk.summary(),
        })
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 200 (10 for --impl reference)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-repeats", type=int, default=5)
    ap.add_argument("--ref-frames", type=int, default=FRAMES_PER_GPU,
                    help="frames of the step the CPU formulation is timed on (default: the whole step)")
    ap.add_argument("--ref-budget-s", type=float, default=150.0,
                    help="--impl reference: wall-clock bound of the whole run; steps shrink to fewer frames beyond it")
    ap.add_argument("--e2e-frames-per-chunk", type=int, default=4)
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        if args.steps is None:
            args.steps = 10
        run_reference(args)
    else:
        if args.steps is None:
            args.steps = 200
        run_ours(args)


if __name__ == "__main__":
    main()
