"""The op for batches that live in HOST memory: forward + backward streamed over the frames.

The reference keeps its frames on the device for the whole step
(/root/reference/models/deformable_transformer.py:191-205), so it has no host entry point; a caller
that holds the tensors of a step in host memory (a data-loader thread, a CPU stage of a pipeline, a
benchmark timing the op end to end) pays the PCIe transfers around the kernels.  Frames are
independent in both passes (/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:263-269),
so the transfers can hide each other and the kernels: the batch is cut into chunks of a few frames
and three streams run a software pipeline

    copy-in stream   H2D(chunk i+1)                       (pinned host -> staging slot)
    compute stream   forward + backward(chunk i)          (the same C-ABI calls as MSDeformAttnFunction)
    copy-out stream  D2H(chunk i-1)                       (results -> pinned host)

PCIe is full duplex, so a step costs about max(H2D, D2H) instead of H2D + kernels + D2H.  Results are
bit-identical to one call over the whole batch (chunking by frames changes no summation order).

No CPU fallback: without a CUDA device (or the library) this raises.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import msda_ext

_SLOTS = 3   # buffer sets in flight: one being filled, one being computed on, one being drained


def chunk_ranges(n_frames: int, frames_per_chunk: int) -> List[Tuple[int, int]]:
    """[start, stop) frame ranges of the pipeline's chunks; the last one may be short."""
    if n_frames < 0 or frames_per_chunk <= 0:
        raise ValueError("n_frames must be >= 0 and frames_per_chunk > 0")
    return [(s, min(s + frames_per_chunk, n_frames)) for s in range(0, n_frames, frames_per_chunk)]


def _require_pinned(named: Sequence[Tuple[str, torch.Tensor]]) -> None:
    for name, t in named:
        if t.is_cuda:
            raise RuntimeError(f"{name} is a CUDA tensor: this entry point takes host buffers "
                               f"(use MSDeformAttnFunction for device tensors)")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
        if not t.is_pinned():
            raise RuntimeError(f"{name} must be in pinned host memory (tensor.pin_memory()) for asynchronous copies")


def ramped_chunk_ranges(n_frames: int, frames_per_chunk: int) -> List[Tuple[int, int]]:
    """Chunks that start and end small (1, 1, 2, 4, ... up to ``frames_per_chunk`` and back down): the first
    copy-in and the last copy-out are the only transfers nothing else can hide, so they are kept short, while
    the chunks in between are large enough to amortise launch and queueing costs."""
    if n_frames < 0 or frames_per_chunk <= 0:
        raise ValueError("n_frames must be >= 0 and frames_per_chunk > 0")
    head, size = [], 1
    while size < frames_per_chunk:
        head.append(size)
        size = size * 2 if len(head) > 1 else 1
    if not head or 2 * sum(head) + frames_per_chunk > n_frames:
        return chunk_ranges(n_frames, frames_per_chunk)
    middle = n_frames - 2 * sum(head)
    sizes = head + [frames_per_chunk] * (middle // frames_per_chunk) + \
        ([middle % frames_per_chunk] if middle % frames_per_chunk else []) + head[::-1]
    out, s0 = [], 0
    for sz in sizes:
        out.append((s0, s0 + sz))
        s0 += sz
    return out


class HostFramePipeline:
    """Reusable pipeline state (streams, staging slots, events) for one device.

        pipe = HostFramePipeline("cuda:0", frames_per_chunk=4)
        out, grad_value, grad_loc, grad_attn = pipe.forward_backward(
            value, spatial_shapes, level_start_index, sampling_locations, attention_weights, grad_output)

    ``graph=True`` replays a captured CUDA graph of the step when it is called again with the same host buffers.

    All six tensors except the two small int64 ones are pinned host tensors with the layouts of
    ``MSDeformAttnFunction``; the four results are pinned host tensors (pass ``results=`` to reuse
    buffers).  The call returns once everything is queued; the results are complete after the current
    stream (which is made to wait for the pipeline) has been synchronised.
    """

    def __init__(self, device="cuda:0", frames_per_chunk: int = 4, im2col_step: int = 64, ramp: bool = True,
                 graph: bool = False):
        if not torch.cuda.is_available():
            raise RuntimeError("HostFramePipeline needs a CUDA device: there is no CPU path for the op")
        self.device = torch.device(device)
        self.frames_per_chunk = int(frames_per_chunk)
        self.im2col_step = im2col_step
        self.ramp = ramp
        # graph=True: the whole step (every copy and launch of every chunk, with the cross-stream ordering) is captured
        # once per set of host buffers and replayed with ONE launch afterwards -- no per-chunk host work, so the copy
        # engines are never left waiting for the next enqueue.  Calls with other buffers capture their own graph.
        self.graph = bool(graph)
        self._graphs: Dict[Tuple, "torch.cuda.CUDAGraph"] = {}
        self.trace = False                     # True: time every stage of the next call with CUDA events
        self._marks: List[Tuple[int, str, torch.cuda.Event, torch.cuda.Event]] = []
        self._t0: Optional[torch.cuda.Event] = None
        with torch.cuda.device(self.device):
            self.s_in, self.s_run, self.s_out = (torch.cuda.Stream() for _ in range(3))
        self._slots: Optional[List[Dict[str, torch.Tensor]]] = None
        self._slot_key = None
        self._computed = [None] * _SLOTS
        self._drained = [None] * _SLOTS
        self._meta: Dict[Tuple, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.launches = 0                      # kernels launched by the last call

    # -- small int64 tables: resident on the device, cached by content
    def _tables(self, spatial_shapes, level_start_index):
        key = (tuple(spatial_shapes.flatten().tolist()), tuple(level_start_index.flatten().tolist()))
        if key not in self._meta:
            self._meta[key] = (spatial_shapes.to(self.device, torch.int64).contiguous(),
                               level_start_index.to(self.device, torch.int64).contiguous())
        return self._meta[key]

    def _staging(self, named):
        """The device side of the pipeline: per slot, the chunk's inputs, its four results, the backward's
        workspace and the forward's index -- allocated once per problem shape, so that a step makes no
        allocator call (a cudaMalloc in the loop would serialise the three streams)."""
        key = tuple((n, tuple(t.shape[1:]), t.dtype) for n, t in named)
        if self._slots is None or self._slot_key != key:
            fpc, dev = self.frames_per_chunk, self.device
            value, loc = named[0][1], named[1][1]
            M, D = value.shape[2], value.shape[3]
            Lq = loc.shape[1]
            self._slots = []
            for _ in range(_SLOTS):
                slot = {n: torch.empty((fpc,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev) for n, t in named}
                slot["out"] = torch.empty((fpc, Lq, M * D), dtype=value.dtype, device=dev)
                slot["gv"] = torch.empty_like(slot["value"])
                slot["gl"] = torch.empty_like(slot["sampling_locations"])
                slot["ga"] = torch.empty_like(slot["attention_weights"])
                slot["ws"] = torch.empty(max(16, msda_ext.backward_workspace_bytes(slot["value"], slot["sampling_locations"])),
                                         dtype=torch.uint8, device=dev)
                slot["index"] = torch.empty(max(16, msda_ext.forward_index_bytes(slot["value"], slot["sampling_locations"])),
                                            dtype=torch.uint8, device=dev)
                self._slots.append(slot)
            self._slot_key = key
            self._computed = [None] * _SLOTS   # event: the kernels that last read the slot's inputs have finished
            self._drained = [None] * _SLOTS    # event: the D2H copies that last read the slot's results have finished
        return self._slots

    def forward_backward(self, value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
                         grad_output, results: Optional[Sequence[torch.Tensor]] = None):
        named = [("value", value), ("sampling_locations", sampling_locations),
                 ("attention_weights", attention_weights), ("grad_output", grad_output)]
        _require_pinned(named)
        N, S, M, D = value.shape
        Lq = sampling_locations.shape[1]
        if grad_output.shape[0] != N or sampling_locations.shape[0] != N or attention_weights.shape[0] != N:
            raise RuntimeError("value, sampling_locations, attention_weights and grad_output must hold the same frames")
        if results is None:
            results = (torch.empty((N, Lq, M * D), dtype=value.dtype).pin_memory(),
                       torch.empty_like(value).pin_memory(),
                       torch.empty_like(sampling_locations).pin_memory(),
                       torch.empty_like(attention_weights).pin_memory())
        else:
            _require_pinned([(f"results[{i}]", r) for i, r in enumerate(results)])
        out_h, gv_h, gl_h, ga_h = results
        shapes_d, lsi_d = self._tables(spatial_shapes, level_start_index)
        slots = self._staging(named)

        with torch.cuda.device(self.device):
            if not self.graph or self.trace:
                self._enqueue(named, results, shapes_d, lsi_d, slots, N)
            else:
                key = tuple(t.data_ptr() for _, t in named) + tuple(r.data_ptr() for r in results) + \
                    (N, self._slot_key, shapes_d.data_ptr(), lsi_d.data_ptr())
                g = self._graphs.get(key)
                if g is None:
                    self._enqueue(named, results, shapes_d, lsi_d, slots, N)      # warm-up (and this call's results)
                    torch.cuda.synchronize(self.device)
                    self._computed, self._drained = [None] * _SLOTS, [None] * _SLOTS   # no events from outside the capture
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=torch.cuda.Stream()):
                        self._enqueue(named, results, shapes_d, lsi_d, slots, N)
                    self._computed, self._drained = [None] * _SLOTS, [None] * _SLOTS   # replays are ordered on the caller's stream
                    self._graphs[key] = g
                else:
                    g.replay()
        return out_h, gv_h, gl_h, ga_h


    def _enqueue(self, named, results, shapes_d, lsi_d, slots, N):
        """Queue one step: every chunk's copy-in, kernels and copy-out on the three streams, ordered by events; the
        caller's (current) stream waits for the last copy-out.  Also what a CUDA graph of the step captures."""
        out_h, gv_h, gl_h, ga_h = results
        caller = torch.cuda.current_stream()
        start = torch.cuda.Event(enable_timing=self.trace)
        start.record(caller)
        self._marks, self._t0 = [], start

        def mark(stream):
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            return e
        self.s_in.wait_event(start)        # host buffers written by work queued on the caller's stream
        self.s_out.wait_event(start)       # ... and result buffers it may still be reading
        self.launches = 0
        last_out = None
        plan = (ramped_chunk_ranges if self.ramp else chunk_ranges)(N, self.frames_per_chunk)
        for i, (lo, hi) in enumerate(plan):
            k = i % _SLOTS
            slot = slots[k]
            n = hi - lo
            if self._computed[k] is not None:
                self.s_in.wait_event(self._computed[k])
            b0 = mark(self.s_in) if self.trace else None
            with torch.cuda.stream(self.s_in):
                for name, t in named:
                    slot[name][:n].copy_(t[lo:hi], non_blocking=True)
            if self.trace:
                self._marks.append((i, "h2d", b0, mark(self.s_in)))
            ready = torch.cuda.Event()
            ready.record(self.s_in)
            self.s_run.wait_event(ready)
            if self._drained[k] is not None:
                self.s_run.wait_event(self._drained[k])
            b0 = mark(self.s_run) if self.trace else None
            with torch.cuda.stream(self.s_run):
                a = (slot["value"][:n], shapes_d, lsi_d, slot["sampling_locations"][:n], slot["attention_weights"][:n])
                _, index = msda_ext.ms_deform_attn_forward(*a, self.im2col_step, want_index=True,
                                                           out=slot["out"][:n], index_buf=slot["index"])   # None for small calls
                self.launches += msda_ext.last_launch_count()
                msda_ext.ms_deform_attn_backward(*a, slot["grad_output"][:n], self.im2col_step, index=index,
                                                 grads=(slot["gv"][:n], slot["gl"][:n], slot["ga"][:n]),
                                                 workspace=slot["ws"])
                self.launches += msda_ext.last_launch_count()
            if self.trace:
                self._marks.append((i, "run", b0, mark(self.s_run)))
            done = torch.cuda.Event()
            done.record(self.s_run)
            self._computed[k] = done
            self.s_out.wait_event(done)
            b0 = mark(self.s_out) if self.trace else None
            with torch.cuda.stream(self.s_out):
                for dst, src in ((out_h, "out"), (gv_h, "gv"), (gl_h, "gl"), (ga_h, "ga")):
                    dst[lo:hi].copy_(slot[src][:n], non_blocking=True)
            if self.trace:
                self._marks.append((i, "d2h", b0, mark(self.s_out)))
            last_out = torch.cuda.Event()
            last_out.record(self.s_out)
            self._drained[k] = last_out
        if last_out is not None:
            caller.wait_event(last_out)    # synchronising the caller's stream now covers the whole pipeline

    def timeline(self) -> List[Tuple[int, str, float, float]]:
        """(chunk, stage, begin ms, end ms) of the last traced call, relative to its start on the caller's
        stream; stages are "h2d", "run" and "d2h".  Synchronises the device."""
        torch.cuda.synchronize(self.device)
        return [(i, st, self._t0.elapsed_time(b), self._t0.elapsed_time(e)) for i, st, b, e in self._marks]


def forward_backward_host(value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
                          grad_output, device="cuda:0", frames_per_chunk: int = 4, im2col_step: int = 64):
    """One-shot convenience wrapper around ``HostFramePipeline`` (synchronises before returning)."""
    pipe = HostFramePipeline(device, frames_per_chunk, im2col_step)
    res = pipe.forward_backward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
                                grad_output)
    torch.cuda.current_stream(pipe.device).synchronize()
    return res
