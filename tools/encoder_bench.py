"""Encoder-level driver (BASELINE.json configs[1]/[2]): three deformable encoder layers -- self-attention
through this repo's MSDeformAttn, FFN 2048, LayerNorms -- over 16 frames x 5100 tokens per GPU, forward +
backward, optionally under bf16 autocast and DistributedDataParallel (gradient all-reduce over NCCL).

The layer restates DeformableTransformerEncoderLayer / DeformableTransformerEncoder
(/root/reference/models/deformable_transformer.py:225-293: self_attn(src + pos) -> residual -> LayerNorm ->
FFN -> residual -> LayerNorm; reference points = pixel centres of every level for every query) with this
repo's module inside; the reference file itself cannot travel to the GPU box.  Dense parts are PyTorch
(cuBLAS); the point is the op inside its real caller: autocast dtype mix, autograd, index handoff, DDP.

    python tools/encoder_bench.py [--layers 3] [--frames 16] [--amp] [--steps 20]
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


def main():
    ap = argparse.ArgumentParser()
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
    model = Encoder(a.layers).to(dev)
    with torch.no_grad():          # leave the all-zero init of the offset / attention projections
        for m in model.modules():
            if isinstance(m, MSDeformAttn):
                m.fused_prologue = a.fuse
                m.sampling_offsets.weight.normal_(0, 0.02)
                m.attention_weights.weight.normal_(0, 0.02)
    net = nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=a.amp):
            out = net(src, pos, ref, shapes, lsi)
            loss = out.float().pow(2).mean()
        loss.backward()
        opt.step()
        return loss

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
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=a.amp):
        out = net(src, pos, ref, shapes, lsi)
    torch.cuda.synchronize()
    fwd_ms = sum(t for _, t in _lib.profile_read())
    _lib.profile_enable(False)
    if rank == 0:
        print(json.dumps({"what": "deformable encoder fwd+bwd+AdamW", "layers": a.layers, "frames_per_gpu": a.frames,
                          "tokens_per_frame": S, "amp_bf16": a.amp, "fused_prologue": a.fuse, "n_gpus": world, "ms_per_step": float(ms.item()),
                          "queries_per_s": world * a.frames * S / (float(ms.item()) * 1e-3),
                          "msda_forward_kernels_ms": fwd_ms, "loss": float(loss)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
