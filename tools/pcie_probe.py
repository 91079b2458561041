"""PCIe probe for the host entry point: H2D alone, D2H alone, both at once, and the frame pipeline at
several chunk sizes (development tool; prints JSON lines)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurips2023_soc_b200.host_frames import HostFramePipeline  # noqa: E402
from neurips2023_soc_b200.synthetic import make_inputs  # noqa: E402


def ev_time(fn, n=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    dev = torch.device("cuda:0")
    nbytes = 208896000
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    t_h2d = ev_time(lambda: d_a.copy_(h_in, non_blocking=True))
    t_d2h = ev_time(lambda: h_out.copy_(d_b, non_blocking=True))

    def both():
        cur = torch.cuda.current_stream()
        e = torch.cuda.Event(); e.record(cur)
        s1.wait_event(e); s2.wait_event(e)
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)
        e1, e2 = torch.cuda.Event(), torch.cuda.Event()
        e1.record(s1); e2.record(s2)
        cur.wait_event(e1); cur.wait_event(e2)
    t_both = ev_time(both)
    print(json.dumps(dict(bytes=nbytes, h2d_ms=t_h2d, d2h_ms=t_d2h, both_ms=t_both, h2d_GBs=nbytes / t_h2d / 1e6,
                          d2h_GBs=nbytes / t_d2h / 1e6, duplex_GBs_each=nbytes / t_both / 1e6)), flush=True)

    host = make_inputs(N=16, dist="encoder", seed=0)
    pin = {k: getattr(host, k).to(torch.bfloat16 if k in ("value", "grad_output") else torch.float32).pin_memory()
           for k in ("value", "sampling_locations", "attention_weights", "grad_output")}
    for fpc, ramp in ((1, False), (2, False), (2, True), (4, False), (4, True), (8, True), (16, False)):
        pipe = HostFramePipeline(dev, frames_per_chunk=fpc, ramp=ramp)
        res = pipe.forward_backward(pin["value"], host.spatial_shapes, host.level_start_index, pin["sampling_locations"],
                                    pin["attention_weights"], pin["grad_output"])
        torch.cuda.synchronize()
        t = ev_time(lambda: pipe.forward_backward(pin["value"], host.spatial_shapes, host.level_start_index,
                                                  pin["sampling_locations"], pin["attention_weights"], pin["grad_output"],
                                                  results=res), n=10)
        import time
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipe.forward_backward(pin["value"], host.spatial_shapes, host.level_start_index, pin["sampling_locations"],
                              pin["attention_weights"], pin["grad_output"], results=res)
        t_queue = (time.perf_counter() - t0) * 1e3
        torch.cuda.synchronize()
        print(json.dumps(dict(frames_per_chunk=fpc, ramp=ramp, ms_per_step=t, launches=pipe.launches, host_queue_ms=t_queue)), flush=True)


if __name__ == "__main__" and "--trace" not in sys.argv:
    main()


def trace():
    dev = torch.device("cuda:0")
    host = make_inputs(N=16, dist="encoder", seed=0)
    pin = {k: getattr(host, k).to(torch.bfloat16 if k in ("value", "grad_output") else torch.float32).pin_memory()
           for k in ("value", "sampling_locations", "attention_weights", "grad_output")}
    pipe = HostFramePipeline(dev, frames_per_chunk=2, ramp=False)
    a = (pin["value"], host.spatial_shapes, host.level_start_index, pin["sampling_locations"], pin["attention_weights"],
         pin["grad_output"])
    res = pipe.forward_backward(*a)
    for _ in range(3):
        pipe.forward_backward(*a, results=res)
    torch.cuda.synchronize()
    pipe.trace = True
    pipe.forward_backward(*a, results=res)
    for i, st, b, e in sorted(pipe.timeline(), key=lambda r: r[2]):
        print(f"chunk {i:2d} {st}  {b:7.3f} -> {e:7.3f}  ({e - b:.3f} ms)")


if __name__ == "__main__" and "--trace" in sys.argv:
    trace()
