"""Frame-sharded inference driver (BASELINE.json configs[3]): a 36-frame video at the 360x640 pyramid, three
deformable encoder layers plus three decoder-style cross-attention layers with 5 queries per frame, forward only.
Frames are dealt out contiguously to the ranks (neurips2023_soc_b200.frames), every rank runs its frames with no
communication, and the per-frame decoder states are gathered once at the end over NCCL -- the reference instead
gives whole videos to processes (/root/reference/infer_refytb.py:92-109) and never shards one video.

By default the transformer IS the reference's: DeformableTransformer of the unmodified models/deformable_transformer.py
staged under baseline/_ref/soc (tools/stage_reference.py) -- 3 encoder + 3 decoder layers, 5 queries per frame, its own
MSDeformAttn module and autograd function on this repo's two extension entry points.  --restated (or nothing staged)
uses the small restatement below.  With --graph (restated model only: the reference's forward reads the level shapes
on the host) the rank's whole forward is captured in a CUDA graph and replayed (the op never synchronises).

    python tools/inference_bench.py [--graph] [--amp]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/inference_bench.py --amp
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurips2023_soc_b200 import MSDeformAttn  # noqa: E402
from neurips2023_soc_b200.frames import frame_range, gather_frames  # noqa: E402
from neurips2023_soc_b200.synthetic import A2D_PYRAMID, level_start_index, pyramid_reference_points  # noqa: E402
from tools.encoder_bench import EncoderLayer  # noqa: E402


class DecoderLayer(nn.Module):
    """nn.MultiheadAttention self-attention over the queries, MSDeformAttn cross-attention into the encoder
    memory, FFN (/root/reference/models/deformable_transformer.py:330-347)."""

    def __init__(self, d_model=256, d_ffn=2048):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, 8, batch_first=True)
        self.cross_attn = MSDeformAttn(d_model, 4, 8, 4)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(d_model), nn.LayerNorm(d_model), nn.LayerNorm(d_model)
        self.linear1, self.linear2 = nn.Linear(d_model, d_ffn), nn.Linear(d_ffn, d_model)

    def forward(self, tgt, query_pos, ref, memory, shapes, lsi):
        q = tgt + query_pos
        tgt = self.norm2(tgt + self.self_attn(q, q, tgt)[0])
        tgt2, _, _ = self.cross_attn(tgt + query_pos, ref, memory, shapes, lsi, None)
        tgt = self.norm1(tgt + tgt2)
        return self.norm3(tgt + self.linear2(torch.relu(self.linear1(tgt))))


class Model(nn.Module):
    def __init__(self):
        super().__init__()
        self.enc = nn.ModuleList(EncoderLayer() for _ in range(3))
        self.dec = nn.ModuleList(DecoderLayer() for _ in range(3))

    def forward(self, src, pos, enc_ref, tgt, query_pos, dec_ref, shapes, lsi):
        for layer in self.enc:
            src = layer(src, pos, enc_ref, shapes, lsi)
        for layer in self.dec:
            tgt = layer(tgt, query_pos, dec_ref, src, shapes, lsi)
        return tgt


def reference_transformer(queries):
    """DeformableTransformer of the staged, unmodified reference file (None when nothing is staged)."""
    import importlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    staged = os.path.join(root, "baseline", "_ref", "soc")
    if not os.path.exists(os.path.join(staged, "MANIFEST.json")):
        return None
    sys.path.insert(1, staged)                 # behind the repo root: MultiScaleDeformableAttention is this repo's shim
    dt = importlib.import_module("models.deformable_transformer")
    return dt.DeformableTransformer(d_model=256, nhead=8, num_encoder_layers=3, num_decoder_layers=3, dim_feedforward=2048,
                                    dropout=0.0, activation="relu", return_intermediate_dec=True, num_feature_levels=4,
                                    dec_n_points=4, enc_n_points=4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--restated", action="store_true", help="the small restatement instead of the staged reference transformer")
    ap.add_argument("--frames", type=int, default=36)
    ap.add_argument("--queries", type=int, default=5)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--amp", action="store_true")
    ap.add_argument("--graph", action="store_true")
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    shapes_l = list(A2D_PYRAMID)
    S, L = sum(h * w for h, w in shapes_l), len(shapes_l)
    shapes = torch.tensor(shapes_l, dtype=torch.long, device=dev)
    lsi = torch.tensor(level_start_index(shapes_l), dtype=torch.long, device=dev)
    lo, hi = frame_range(a.frames, world, rank)
    n = hi - lo
    g = torch.Generator().manual_seed(0)                  # the same video on every rank; each keeps its frames
    src = torch.randn(a.frames, S, 256, generator=g)[lo:hi].to(dev)
    pos = torch.randn(a.frames, S, 256, generator=g)[lo:hi].to(dev)
    tgt = torch.randn(a.frames, a.queries, 256, generator=g)[lo:hi].to(dev)
    qpos = torch.randn(a.frames, a.queries, 256, generator=g)[lo:hi].to(dev)
    dec_ref = torch.rand(a.frames, a.queries, 1, 2, generator=g)[lo:hi].to(dev).expand(n, a.queries, L, 2).contiguous()
    enc_ref = pyramid_reference_points(shapes_l).to(dev)[None, :, None, :].expand(n, S, L, 2).contiguous()
    torch.manual_seed(0)
    ref_model = None if (a.restated or a.graph) else reference_transformer(a.queries)
    which = "reference DeformableTransformer (staged, unmodified)" if ref_model is not None else "restated layers"
    model = (ref_model if ref_model is not None else Model()).to(dev).eval()
    with torch.no_grad():
        for m in model.modules():
            if hasattr(m, "sampling_offsets") and hasattr(m, "attention_weights"):
                m.sampling_offsets.weight.normal_(0, 0.02)
                m.attention_weights.weight.normal_(0, 0.02)
    if ref_model is not None:      # the reference's calling convention: per-level maps, masks, positions; tgt [b, t, q, c]
        srcs = [src[:, s0:s0 + h * w].transpose(1, 2).reshape(n, 256, h, w).contiguous()
                for (h, w), s0 in zip(shapes_l, level_start_index(shapes_l))]
        poses = [pos[:, s0:s0 + h * w].transpose(1, 2).reshape(n, 256, h, w).contiguous()
                 for (h, w), s0 in zip(shapes_l, level_start_index(shapes_l))]
        masks = [torch.zeros(n, h, w, dtype=torch.bool, device=dev) for h, w in shapes_l]
        query_embed = qpos[0].contiguous()

    def fwd():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=a.amp):
            if ref_model is not None:
                return model(srcs, tgt[None], masks, poses, query_embed)[0][-1]      # last decoder layer: [frames, q, c]
            return model(src, pos, enc_ref, tgt, qpos, dec_ref, shapes, lsi)

    for _ in range(3):
        out = fwd()
    torch.cuda.synchronize()
    graph = None
    if a.graph:
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.cuda.graph(graph, stream=side):
            out = fwd()
        torch.cuda.current_stream().wait_stream(side)
        graph.replay()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        if graph is not None:
            graph.replay()
        else:
            out = fwd()
        full = gather_frames(out.float(), a.frames) if world > 1 else out.float()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"what": "36-frame video: 3 encoder + 3 decoder layers, forward, frames sharded, output gathered",
                          "model": which,
                          "frames": a.frames, "queries_per_frame": a.queries, "n_gpus": world, "amp_bf16": a.amp,
                          "cuda_graph": a.graph, "ms_per_video": float(ms.item()),
                          "frames_per_s": a.frames / (float(ms.item()) * 1e-3),
                          "gathered_shape": list(full.shape), "checksum": float(full.double().sum())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
