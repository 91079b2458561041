#!/usr/bin/env python
"""bench.py -- MSDeformAttn forward + backward at the A2D-Sentences Video-Swin-T encoder shape.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 8-frame clips, batch 2 => 16 frames per GPU, 5100 tokens per
frame (48x80 / 24x40 / 12x20 / 6x10), d_model 256 = 8 heads x 32, 4 levels x 4 points, queries =
tokens (encoder self-attention).  bf16 values / outputs / value gradients, fp32 sampling locations
and attention weights (what autocast produces), fp32 accumulation.  One STEP = one forward + one
backward of the op over the rank's frames.  Frames are independent, so ranks hold disjoint frames and
there is no data-path collective; `value` = queries of all ranks / time (max over ranks).
  --scaling weak (default): 16 frames per rank.  --scaling strong: 16 frames in total, 16 / N per rank.
  A weak run on N > 1 ranks also times the strong split and reports it under "strong_scaling".

The timed step uses buffers allocated once and, by default, a CUDA graph of the step's launches
(--no-graph: eager calls through ctypes), so that the host is not on the critical path.

Printed JSON line (rank 0): metric/value/unit..., plus
  roofline     -- the dominant kernel (longest device time per step, measured live with CUDA events
                  through msda_profile_*): its ALGORITHMIC bytes per launch (tensors of the op
                  only -- the inverse index the backward builds for itself is reported apart as
                  `index_bytes_per_launch`) over its average duration, against MEASURED_PEAKS.json's copy
                  bandwidth; `step` does the same for the whole forward + backward; `per_kernel` lists every
                  kernel; `l1_wavefront` is the gather ceiling of DESIGN.md section 5, clearly not the HBM one
  ref_cuda     -- the reference's own CUDA kernels recompiled for sm_100a (oracle/_ref, when shipped) timed on
                  the same fp32 tensors beside this repo's kernels: the GPU bar to beat
  unordered    -- the same step with MSDA_FLAG_UNORDERED (opt-in: no rank sort, grad_value in arrival order like the
                  reference's atomics): what bit-reproducibility costs; `value` is always the reproducible default
  cpu_baseline -- the reference's CPU formulation (ms_deform_attn_core_pytorch: the staged reference file
                  when baseline/_ref/soc exists, else its restatement in oracle/) timed on this box's host cores
  e2e          -- the same step through the host entry point (host_frames.HostFramePipeline) with inputs in
                  pinned HOST memory: H2D of value/locations/weights/grad_output and D2H of output + the
                  three gradients are inside the timed region, pipelined over chunks of frames
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


def base_config(world: int, scaling: str) -> dict:
    """The workload description both arms print (the driver compares the two `config` objects)."""
    per = FRAMES_PER_GPU if scaling == "weak" else FRAMES_PER_GPU // world
    return {"workload": WORKLOAD, "frames_per_gpu": per, "queries_per_step_per_gpu": per * 5100,
            "value_dtype": "bf16", "location_weight_dtype": "f32", "accumulate": "f32",
            "locations": "encoder-realistic (SURVEY.md 8d distribution A)",
            "l2": "per-step working set ~0.7 GB (inputs 0.25 GB + outputs/gradients 0.25 GB + index 0.2 GB) "
                  "exceeds the 126 MB L2; no explicit flush",
            "parallelism": f"frames sharded over {world} rank(s), no data-path collective"}


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
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs (rank 0 only: one
    sampler per rank was eight extra processes on the host of an 8-GPU run)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, enabled: bool = True):
        self.index, self.rows, self.proc, self.enabled = index, [], None, enabled

    def __enter__(self):
        if not self.enabled:
            return self
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
    """(callable, kind): the reference's CPU path, ms_deform_attn_core_pytorch
    (/root/reference/models/ops/functions/ms_deform_attn_func.py:41-61).  kind "reference": the reference's own
    file, staged byte for byte under baseline/_ref/soc by tools/stage_reference.py (its
    `import MultiScaleDeformableAttention` binds to this repo's shim, which only loads the library when called);
    kind "port": the restatement in oracle/, pinned to it by tests/test_oracle_golden.py.  All host threads."""
    torch.set_num_threads(threads)
    staged = ROOT / "baseline" / "_ref" / "soc" / "models" / "ops" / "functions" / "ms_deform_attn_func.py"
    if staged.exists():
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("soc_reference_ms_deform_attn_func", staged)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            core = mod.ms_deform_attn_core_pytorch
            return (lambda v, shapes, lo, at: core(v, shapes, lo, at)), "reference"
        except Exception as e:      # fall back to the restatement, say so on stderr
            print(f"[bench] staged reference function not importable ({e}); using the oracle port", file=sys.stderr)
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
        out.backward(x.grad_output.view_as(out))
        dt = time.perf_counter() - t0
        if i or not warm:          # with warm, the first pass is the warm-up
            best = min(best, dt)
    return best


def run_reference(args):
    from neurips2023_soc_b200.synthetic import make_inputs
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
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
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(max(world, args.gpus, 1), args.scaling),
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# --------------------------------------------------------------------------------------------
class DeviceStep:
    """One forward + backward of the op over `x` with every buffer allocated once (output, the three gradients,
    the backward's workspace, the forward's index) -- what MSDeformAttnFunction does minus the allocator calls --
    optionally captured in a CUDA graph."""

    def __init__(self, x, msda_ext, graph: bool, flags=None):
        self.x, self.ext, self.flags = x, msda_ext, flags
        self.args = (x.value, x.spatial_shapes, x.level_start_index, x.sampling_locations, x.attention_weights)
        N, S, M, D = x.value.shape
        Lq = x.sampling_locations.shape[1]
        dev = x.value.device
        self.out = torch.empty((N, Lq, M * D), dtype=x.value.dtype, device=dev)
        self.grads = (torch.empty_like(x.value), torch.empty_like(x.sampling_locations), torch.empty_like(x.attention_weights))
        self.ws = torch.empty(max(16, msda_ext.backward_workspace_bytes(x.value, x.sampling_locations)), dtype=torch.uint8, device=dev)
        self.index = torch.empty(max(16, msda_ext.forward_index_bytes(x.value, x.sampling_locations)), dtype=torch.uint8, device=dev)
        self.launches = 0
        self.graph = None
        self.mode = "eager (ctypes calls, buffers allocated once)"
        for _ in range(3):
            self.eager()
        torch.cuda.synchronize()
        if graph:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.eager()
                self.graph = g
                self.mode = "cuda_graph (one graph launch per step, buffers allocated once)"
            except Exception as e:
                print(f"[bench] CUDA graph capture failed ({e}); timing eager calls", file=sys.stderr)
                torch.cuda.synchronize()

    def eager(self):
        _, index = self.ext.ms_deform_attn_forward(*self.args, 64, want_index=True, out=self.out, index_buf=self.index)
        lf = self.ext.last_launch_count()
        self.ext.ms_deform_attn_backward(*self.args, self.x.grad_output, 64, index=index, grads=self.grads, workspace=self.ws,
                                         flags=self.flags)
        self.launches = lf + self.ext.last_launch_count()

    def __call__(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.eager()


def run_ours(args):
    import torch.distributed as dist
    from neurips2023_soc_b200 import MSDeformAttnFunction, _lib, msda_ext
    from neurips2023_soc_b200.synthetic import algorithmic_bytes, make_inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the op (use --impl reference)")
    if args.scaling == "strong" and FRAMES_PER_GPU % world:
        raise SystemExit(f"--scaling strong splits {FRAMES_PER_GPU} frames: the rank count must divide it")
    # each rank on its own slice of the host's cores: the e2e leg is host-memory / PCIe work
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(1, world))
        mine = cores[(local * per) % len(cores):(local * per) % len(cores) + per]
        if world > 1 and mine:
            os.sched_setaffinity(0, mine)
    except Exception:
        pass
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    vdt, adt = torch.bfloat16, torch.float32
    frames = FRAMES_PER_GPU if args.scaling == "weak" else FRAMES_PER_GPU // world
    host_all = make_inputs(N=FRAMES_PER_GPU, dist="encoder", seed=rank if args.scaling == "weak" else 0)

    def take(h, lo, hi):          # frames [lo, hi) of a host batch
        import copy
        c = copy.copy(h)
        for k in ("value", "sampling_locations", "attention_weights", "grad_output"):
            setattr(c, k, getattr(h, k)[lo:hi].contiguous())
        return c
    host = host_all if args.scaling == "weak" else take(host_all, rank * frames, (rank + 1) * frames)
    x = host.to(dev, vdt, adt)
    N, S, M, D = x.value.shape
    Lq, L, P = x.sampling_locations.shape[1], x.sampling_locations.shape[3], x.sampling_locations.shape[4]
    queries = N * Lq
    total_queries = world * queries

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def time_steps(fn, steps, warmup):
        for _ in range(max(3, warmup)):
            fn()
        sync_all()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(steps):
            fn()
        ev[1].record()
        sync_all()
        return max_over_ranks(ev[0].elapsed_time(ev[1])) / steps

    step = DeviceStep(x, msda_ext, graph=not args.no_graph)
    launches_per_step = step.launches

    # ---- timed region: K steps, device time on the launching (current) stream, max over ranks ----
    with ClockSampler(local, enabled=(rank == 0)) as clk:
        ms_step = time_steps(step, args.steps, args.warmup)
    value = total_queries / (ms_step * 1e-3)
    eager_ms = None
    if step.graph is not None:    # the same step issued call by call, for the record
        eager_ms = time_steps(step.eager, min(args.steps, 50), 3)

    # ---- what bit-reproducibility costs: the same step with MSDA_FLAG_UNORDERED (no rank sort; grad_value summed in
    # arrival order, the reference's own behaviour).  For the record only: `value` above is the reproducible default.
    unordered = None
    if rank == 0 and world == 1:
        su = DeviceStep(x, msda_ext, graph=not args.no_graph, flags=_lib.FLAG_UNORDERED)
        n_u = min(args.steps, 50)
        for _ in range(3):
            su()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_u):
            su()
        e1.record()
        torch.cuda.synchronize()
        ms_u = e0.elapsed_time(e1) / n_u
        unordered = {"flag": "MSDA_FLAG_UNORDERED", "ms_per_step": ms_u, "value": total_queries / (ms_u * 1e-3), "unit": UNIT,
                     "steps": n_u, "launches_per_step": su.launches,
                     "note": "opt-in; grad_value not bit-reproducible from run to run (like the reference's atomicAdd)"}
        del su

    # ---- strong scaling beside a weak run: 16 frames in total over the same ranks ----
    strong = None
    if args.scaling == "weak" and world > 1 and FRAMES_PER_GPU % world == 0:
        per = FRAMES_PER_GPU // world
        hs_ = take(make_inputs(N=FRAMES_PER_GPU, dist="encoder", seed=0), rank * per, (rank + 1) * per)
        sstep = DeviceStep(hs_.to(dev, vdt, adt), msda_ext, graph=not args.no_graph)
        s_ms = time_steps(sstep, args.steps, args.warmup)
        strong = {"frames_total": FRAMES_PER_GPU, "frames_per_gpu": per, "ms_per_step": s_ms,
                  "value": FRAMES_PER_GPU * Lq / (s_ms * 1e-3), "unit": UNIT, "launch": sstep.mode,
                  "note": "same kernels on 16 / N frames per rank: per-rank work shrinks with N, so the fixed costs of a "
                          "call (launches, the per-(frame, head) scan, tail of the persistent grids) show"}
        del sstep

    # ---- per-kernel device times over K eager steps (instrumented pass: CUDA events around every launch) ----
    _lib.profile_enable(True)
    prof_steps = min(args.steps, 50)
    for _ in range(prof_steps):
        step.eager()
    torch.cuda.synchronize()
    recs = _lib.profile_read()
    _lib.profile_enable(False)
    per_kernel = {}
    for name, ms in recs:
        per_kernel.setdefault(name, []).append(ms)
    kern_avg = {k: sum(v) / len(v) for k, v in per_kernel.items()}
    kern_per_step = {k: sum(v) / prof_steps for k, v in per_kernel.items()}
    dominant = max(kern_per_step, key=kern_per_step.get)
    vb, ab = x.value.element_size(), x.sampling_locations.element_size()
    samples = N * Lq * M * L * P
    C = M * D
    fwd_bytes, bwd_bytes = algorithmic_bytes(N, S, M, D, L, Lq, P, vb, ab)
    # SURVEY.md 8d bytes of each kernel's own job, per launch: tensors of the op only.  The 16-byte-per-sample
    # inverse index (this repo's device for a deterministic grad_value) is not algorithmic: listed apart.
    alg = {
        "msda_fwd_tile_kernel": fwd_bytes,
        "msda_bwd_sample_tile_kernel": vb * N * S * C + vb * N * Lq * C + ab * samples * 3 + ab * samples * 3,
        "msda_grad_value_walk_kernel": vb * N * Lq * C + vb * N * S * C,
        "msda_bwd_bin_kernel": vb * N * Lq * C + vb * N * S * C,
    }
    idx = {
        "msda_bwd_sample_tile_kernel": 16 * samples,            # entries written
        "msda_bin_rank_sort_kernel": 2 * 16 * samples,          # read + rewritten
        "msda_grad_value_walk_kernel": 16 * samples,            # read
        "msda_bwd_bin_kernel": 16 * samples,
    }
    peak, peak_src = peaks()
    per_kernel_roofline = {}
    for k, ms in sorted(kern_per_step.items(), key=lambda kv: -kv[1]):
        a_bytes = alg.get(k, 0)
        per_kernel_roofline[k] = {"ms_per_step": round(ms, 4), "algorithmic_bytes_per_launch": a_bytes,
                                  "index_bytes_per_launch": idx.get(k, 0),
                                  "frac": (a_bytes / (kern_avg[k] * 1e-3) / 1e9 / peak) if a_bytes else 0.0}
    dom_bytes = alg.get(dominant, 0)
    achieved = dom_bytes / (kern_avg[dominant] * 1e-3) / 1e9
    device_ms = sum(kern_per_step.values())
    step_bytes = fwd_bytes + bwd_bytes
    # the gather ceiling: one L1 wavefront per gathered row, one wavefront per clock and SM
    rows = N * Lq * M * L * P * 4 * 2 + N * Lq * M * L * P      # value rows in forward and part A, grad_output rows in part B
    sm_clock_hz = 1.965e9
    l1_floor_ms = rows / 148.0 / sm_clock_hz * 1e3
    roofline = {
        "bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": dom_bytes, "index_bytes_per_launch": idx.get(dominant, 0),
        "avg_launch_ms": kern_avg[dominant],
        "kernel_ms_per_step": {k: round(v, 4) for k, v in sorted(kern_per_step.items(), key=lambda kv: -kv[1])},
        "per_kernel": per_kernel_roofline,
        "step": {"algorithmic_bytes": step_bytes, "device_ms": device_ms, "timed_ms": ms_step,
                 "achieved": step_bytes / (ms_step * 1e-3) / 1e9,
                 "frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak,
                 "frac_of_nominal_8TBs": step_bytes / (ms_step * 1e-3) / 1e9 / 8000.0,
                 "index_bytes": sum(idx.get(k, 0) for k in kern_per_step)},
        "l1_wavefront": {"what": "NOT the HBM roofline: the floor a row-granular bilinear gather has on this memory hierarchy "
                                 "(one L1 wavefront per 64-byte row, one wavefront per clock and SM at 1965 MHz, 148 SMs)",
                         "rows_gathered_per_step": rows, "floor_ms": l1_floor_ms, "frac": l1_floor_ms / ms_step},
    }
    traffic_file = ROOT / "profiles" / "dominant_kernel_traffic.json"
    if traffic_file.exists():
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(dominant)
        except Exception:
            pass

    # ---- the GPU bar to beat: the reference's own CUDA kernels (sm_100a build) on the same tensors, fp32 ----
    ref_cuda = None
    if rank == 0 and not args.no_ref_cuda:
        try:
            from oracle import build_ref
            ref = build_ref.load()
        except Exception as e:
            ref = None
            print(f"[bench] oracle/_ref not loadable: {e}", file=sys.stderr)
        if ref is not None:
            x32 = host.to(dev, torch.float32, torch.float32)
            a32 = (x32.value, x32.spatial_shapes, x32.level_start_index, x32.sampling_locations, x32.attention_weights)

            def ev_us(fn, iters):
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) * 1e3 / iters
            it = max(3, min(20, args.steps))
            r_f = ev_us(lambda: ref.ms_deform_attn_forward(*a32, 64), it)
            r_b = ev_us(lambda: ref.ms_deform_attn_backward(*a32, x32.grad_output, 64), it)
            s32 = DeviceStep(x32, msda_ext, graph=False)
            o_f = ev_us(lambda: msda_ext.ms_deform_attn_forward(*a32, 64, out=s32.out), it)
            o_fb = ev_us(s32.eager, it)
            ref_cuda = {
                "what": "reference CUDA sources (ms_deform_im2col_cuda.cuh) compiled untouched for sm_100a, same fp32 "
                        "tensors, CUDA events, back to back, no L2 flush",
                "frames": N, "iters": it,
                "reference_fwd_us": r_f, "reference_bwd_us": r_b, "reference_fwd_bwd_us": r_f + r_b,
                "ours_fp32_fwd_us": o_f, "ours_fp32_fwd_bwd_us": o_fb,
                "ours_bf16mix_fwd_bwd_us": ms_step * 1e3,
                "speedup_fp32_fwd": r_f / o_f, "speedup_fp32_fwd_bwd": (r_f + r_b) / o_fb,
                "speedup_bf16mix_fwd_bwd": (r_f + r_b) / (ms_step * 1e3),
            }
            del s32, x32

    # ---- end to end through the public API with host buffers ----
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
    pipe = HostFramePipeline(dev, frames_per_chunk=min(args.e2e_frames_per_chunk, N), graph=args.e2e_graph)
    results = (res_host["out"], res_host["gv"], res_host["gl"], res_host["ga"])

    def e2e_step():
        pipe.forward_backward(pin["value"], host.spatial_shapes, host.level_start_index, pin["sampling_locations"],
                              pin["attention_weights"], pin["grad_output"], results=results)

    e2e_steps = max(2, min(args.steps, 20))
    e2e_ms = time_steps(e2e_step, e2e_steps, 2)
    e2e_serial_ms = time_steps(e2e_unpipelined, max(2, min(args.steps, 5)), 2)

    # what the host can move at all: the step's bytes as plain pinned copies, both directions at once, all ranks at once
    d_in = torch.empty(h2d, dtype=torch.uint8, device=dev)
    d_outb = torch.empty(d2h, dtype=torch.uint8, device=dev)
    h_in = torch.empty(h2d, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(d2h, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def duplex_copy():
        cur = torch.cuda.current_stream()
        e = torch.cuda.Event()
        e.record(cur)
        s1.wait_event(e)
        s2.wait_event(e)
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_outb, non_blocking=True)
        e1, e2 = torch.cuda.Event(), torch.cuda.Event()
        e1.record(s1)
        e2.record(s2)
        cur.wait_event(e1)
        cur.wait_event(e2)
    copy_ms = time_steps(duplex_copy, 5, 2)
    del d_in, d_outb, h_in, h_out
    e2e = {"value": total_queries / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps,
           "api": f"HostFramePipeline.forward_backward: pinned host buffers, chunks of up to {pipe.frames_per_chunk} "
                  f"frames (short first and last chunks), H2D / kernels / D2H pipelined on three streams ({pipe.launches} kernel launches per step)"
                  + (", the step replayed as one CUDA graph" if pipe.graph else ", enqueued call by call"),
           "one_stream_ms_per_step": e2e_serial_ms,
           "pcie_GBs_each_way": max(h2d, d2h) / (e2e_ms * 1e-3) / 1e9,
           "host_copy_ceiling": {"what": "the step's bytes as bare pinned copies, H2D and D2H at once, all ranks at once "
                                         "(max over ranks): the bound the host's memory and PCIe roots put on e2e",
                                 "ms_per_step": copy_ms, "GBs_each_way": max(h2d, d2h) / (copy_ms * 1e-3) / 1e9,
                                 "e2e_over_ceiling": copy_ms / e2e_ms}}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the same workload on the host cores ----
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
        cfg = base_config(world, args.scaling)
        cfg["launch"] = step.mode
        emit({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": cfg,
            "eager_ms_per_step": eager_ms,
            "roofline": roofline, "ref_cuda": ref_cuda, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "strong_scaling": strong, "unordered": unordered,
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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 16 frames per rank; strong: 16 frames in total, 16 / N per rank")
    ap.add_argument("--no-graph", action="store_true", help="time eager ctypes calls instead of a CUDA graph of the step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference-CUDA leg")
    ap.add_argument("--cpu-repeats", type=int, default=5)
    ap.add_argument("--ref-frames", type=int, default=FRAMES_PER_GPU,
                    help="frames of the step the CPU formulation is timed on (default: the whole step)")
    ap.add_argument("--ref-budget-s", type=float, default=150.0,
                    help="--impl reference: wall-clock bound of the whole run; steps shrink to fewer frames beyond it")
    ap.add_argument("--e2e-frames-per-chunk", type=int, default=4)
    ap.add_argument("--e2e-graph", action="store_true",
                    help="replay the host pipeline's step as one CUDA graph (measured: no faster -- the copies bound it)")
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
