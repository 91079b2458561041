"""Encoder-level driver (BASELINE.json configs[1]/[2]): three deformable encoder layers -- self-attention
through this repo's MSDeformAttn, FFN 2048, LayerNorms -- over 16 frames x 5100 tokens per GPU, forward +
backward, optionally under bf16 autocast and DistributedDataParallel (gradient all-reduce over NCCL).

By default the encoder IS the reference's: DeformableTransformerEncoder / DeformableTransformerEncoderLayer of the
unmodified models/deformable_transformer.py staged under baseline/_ref/soc (tools/stage_reference.py), with
`models.ops.modules.MSDeformAttn` routed to this repo's module (--reference-module keeps the reference's own module
and autograd function too, on this repo's two extension entry points).  --restated (or nothing staged) uses the
small restatement below (/root/reference/models/deformable_transformer.py:225-293: self_attn(src + pos) ->
residual -> LayerNorm -> FFN -> residual -> LayerNorm; reference points = pixel centres of every level for every
query).  Dense parts are PyTorch (cuBLAS); the point is the op inside its real caller: autocast dtype mix,
autograd, index handoff, DDP.

    python tools/encoder_bench.py [--layers 3] [--frames 16] [--amp] [--steps 20] [--profile]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/encoder_bench.py --amp
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurips2023_soc_b200 import MSDeformAttn, _lib  # noqa: E402
from neurips2023_soc_b200.synthetic import A2D_PYRAMID, level_start_index, pyramid_reference_points  # noqa: E402


class EncoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=2048, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1, self.linear2 = nn.Linear(d_model, d_ffn), nn.Linear(d_ffn, d_model)
        self.norm2 = nn.LayerNorm(d_model)

    def forward(self, src, pos, reference_points, shapes, lsi):
        src2, _, _ = self.self_attn(src + pos, reference_points, src, shapes, lsi, None)
        src = self.norm1(src + src2)
        return self.norm2(src + self.linear2(torch.relu(self.linear1(src))))


class Encoder(nn.Module):
    def __init__(self, layers):
        super().__init__()
        self.layers = nn.ModuleList(EncoderLayer() for _ in range(layers))

    def forward(self, src, pos, reference_points, shapes, lsi):
        for layer in self.layers:
            src = layer(src, pos, reference_points, shapes, lsi)
        return src


def reference_encoder(layers, keep_reference_module):
    """DeformableTransformerEncoder of the staged, unmodified reference file (None when nothing is staged)."""
    import importlib
    import types
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    staged = os.path.join(root, "baseline", "_ref", "soc")
    if not os.path.exists(os.path.join(staged, "MANIFEST.json")):
        return None
    sys.path.insert(1, staged)                 # behind the repo root: MultiScaleDeformableAttention is this repo's shim
    if not keep_reference_module:              # deformable_transformer.py:20 `from models.ops.modules import MSDeformAttn`
        ops = types.ModuleType("models.ops")
        ops.__path__ = []
        mods = types.ModuleType("models.ops.modules")
        mods.MSDeformAttn = MSDeformAttn
        sys.modules["models.ops"], sys.modules["models.ops.modules"] = ops, mods
    dt = importlib.import_module("models.deformable_transformer")
    layer = dt.DeformableTransformerEncoderLayer(256, 2048, 0.0, "relu", 4, 8, 4)
    return dt.DeformableTransformerEncoder(layer, layers)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--restated", action="store_true", help="the small restatement of the encoder instead of the staged reference class")
    ap.add_argument("--b200-layers", action="store_true",
                    help="this repo's DeformableTransformerEncoder of fused layers (bf16 end to end, SURVEY.md 8f-2)")
    ap.add_argument("--reference-module", action="store_true", help="keep the reference's own MSDeformAttn module / autograd function too")
    ap.add_argument("--graph", action="store_true", help="capture the whole training step (forward, backward, AdamW) in a CUDA graph")
    ap.add_argument("--profile", action="store_true", help="print the step's CUDA kernels by total time (torch.profiler)")
    ap.add_argument("--pad", action="store_true",
                    help="pass a padding mask (all False, as for A2D's equal-size frames: models/soc.py always passes one)")
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--amp", action="store_true")
    ap.add_argument("--fuse", action="store_true", help="softmax / location arithmetic inside the forward kernel")
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(rank)
    shapes_l = list(A2D_PYRAMID)
    S = sum(h * w for h, w in shapes_l)
    shapes = torch.tensor(shapes_l, dtype=torch.long, device=dev)
    lsi = torch.tensor(level_start_index(shapes_l), dtype=torch.long, device=dev)
    ref = pyramid_reference_points(shapes_l).to(dev)[None, :, None, :].expand(a.frames, S, len(shapes_l), 2).contiguous()
    src = torch.randn(a.frames, S, 256, device=dev)
    pos = torch.randn(a.frames, S, 256, device=dev)
    padmask = torch.zeros(a.frames, S, dtype=torch.bool, device=dev) if a.pad else None
    model = None if (a.restated or a.b200_layers) else reference_encoder(a.layers, a.reference_module)
    which = "reference DeformableTransformerEncoder (staged, unmodified)" + (" + reference MSDeformAttn module" if a.reference_module else "")
    if a.b200_layers:
        from neurips2023_soc_b200 import DeformableTransformerEncoder, DeformableTransformerEncoderLayer
        model = DeformableTransformerEncoder(DeformableTransformerEncoderLayer(256, 2048, 0.0, "relu", 4, 8, 4), a.layers)
        which = "neurips2023_soc_b200.DeformableTransformerEncoder (fused layers, bf16 end to end)"
        ratios = torch.ones(a.frames, len(shapes_l), 2, device=dev)
        call = lambda net: net(src, shapes, lsi, ratios, pos, padmask)                        # noqa: E731
    elif model is None:
        model, which = Encoder(a.layers), "restated encoder"
        call = lambda net: net(src, pos, ref, shapes, lsi)                                    # noqa: E731
    else:
        ratios = torch.ones(a.frames, len(shapes_l), 2, device=dev)                           # no padding: valid ratio 1
        call = lambda net: net(src, shapes, lsi, ratios, pos, padmask)                        # noqa: E731
    model = model.to(dev)
    with torch.no_grad():          # leave the all-zero init of the offset / attention projections
        for m in model.modules():
            if hasattr(m, "sampling_offsets") and hasattr(m, "attention_weights"):
                if isinstance(m, MSDeformAttn):
                    m.fused_prologue = a.fuse
                m.sampling_offsets.weight.normal_(0, 0.02)
                m.attention_weights.weight.normal_(0, 0.02)
    net = nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, capturable=a.graph)

    def eager_step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=a.amp):
            out = call(net)
            loss = out.float().pow(2).mean()
        loss.backward()
        opt.step()
        return loss

    step = eager_step
    if a.graph:
        # whole-step capture (torch's recipe): warm up on a side stream, then record forward, backward and the
        # optimizer step once; the op never synchronises and every buffer it allocates comes from the graph's pool
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                eager_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            static_loss = eager_step()

        def step():
            graph.replay()
            return static_loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # share of the step spent in this repo's kernels (main thread = forward; backward runs on autograd's thread)
    _lib.profile_enable(True)
    step()
    torch.cuda.synchronize()
    fwd_ms = sum(t for _, t in _lib.profile_read())      # process-wide: forward and backward kernels of one step
    _lib.profile_enable(False)
    if a.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90), file=sys.stderr)
    if rank == 0:
        print(json.dumps({"what": "deformable encoder fwd+bwd+AdamW", "encoder": which, "layers": a.layers, "frames_per_gpu": a.frames,
                          "tokens_per_frame": S, "amp_bf16": a.amp, "padding_mask": bool(a.pad), "fused_prologue": a.fuse, "cuda_graph": a.graph, "n_gpus": world, "ms_per_step": float(ms.item()),
                          "queries_per_s": world * a.frames * S / (float(ms.item()) * 1e-3),
                          "msda_kernels_ms_per_step": fwd_ms, "loss": float(loss.detach())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
