"""Frame sharding of the hot path over the GPUs of one node.

Every frame ``n`` of the batch dimension (= B*T, ordered ``(b t)``,
/root/reference/models/deformable_transformer.py:191) is independent in both passes of the op:
value, locations, outputs and all three gradients are indexed by ``n`` first
(/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:263-269).  So the op shards with no
exchange at all: each rank takes a contiguous block of frames (whole clips when ``clip_len`` is
given).  The only collectives are the two the reference system has around the path:

* training  -- the gradient all-reduce of the trainable parameters (DDP in the reference,
  /root/reference/trainer.py:52-54): ``allreduce_gradients`` for modules used outside DDP;
* inference -- one gather of the per-frame outputs at the end (the reference instead splits
  videos over processes, /root/reference/infer_refytb.py:92-109): ``gather_frames``.

One process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def frame_range(n_frames: int, world: int, rank: int, clip_len: int = 1) -> Tuple[int, int]:
    """[start, stop) of the frames rank ``rank`` owns: clips (``clip_len`` consecutive frames) are
    dealt out contiguously, the first ``n_clips % world`` ranks get one clip more."""
    if n_frames % clip_len:
        raise ValueError(f"{n_frames} frames do not split into clips of {clip_len}")
    clips = n_frames // clip_len
    base, extra = divmod(clips, world)
    start = rank * base + min(rank, extra)
    stop = start + base + (1 if rank < extra else 0)
    return start * clip_len, stop * clip_len


def shard_frames(tensors: Sequence[torch.Tensor], world: int, rank: int, clip_len: int = 1) -> List[torch.Tensor]:
    """Slice dim 0 (frames) of every tensor to this rank's block; views, no copies."""
    lo, hi = frame_range(tensors[0].shape[0], world, rank, clip_len)
    return [t[lo:hi] for t in tensors]


def gather_frames(local: torch.Tensor, n_frames: int, clip_len: int = 1,
                  group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All ranks' frame blocks concatenated along dim 0 (inverse of ``shard_frames``).  Blocks may
    differ by one clip; they are padded to the largest for the collective and trimmed after."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [hi - lo for lo, hi in (frame_range(n_frames, world, r, clip_len) for r in range(world))]
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} frames, expected {sizes[rank]}")
    big = max(sizes)
    buf = local if local.shape[0] == big else torch.cat(
        [local, local.new_zeros((big - local.shape[0],) + tuple(local.shape[1:]))], 0)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf.contiguous(), group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                        bucket_bytes: int = 64 << 20, average: bool = True) -> int:
    """Sum (or average) ``p.grad`` over the ranks in flat buckets; returns the number of buckets.
    The reference trains 49.5 M parameters = 198 MB of fp32 gradients; NVSwitch gives
    every GPU full bandwidth to every peer, so buckets are sized for launch latency, not links."""
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    buckets, cur, cur_bytes = [], [], 0
    for g in grads:
        nbytes = g.numel() * g.element_size()
        if cur and (cur_bytes + nbytes > bucket_bytes or g.dtype != cur[0].dtype):
            buckets.append(cur)
            cur, cur_bytes = [], 0
        cur.append(g)
        cur_bytes += nbytes
    if cur:
        buckets.append(cur)
    for b in buckets:
        flat = torch.cat([g.reshape(-1) for g in b])
        dist.all_reduce(flat, group=group)
        if average:
            flat /= world
        off = 0
        for g in b:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
    return len(buckets)
