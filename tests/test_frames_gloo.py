"""Frame sharding over ranks: bookkeeping plus the two collectives around the hot path, on CPU with
the gloo backend and world_size 2 (the GPU runs use the same code over NCCL)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neurips2023_soc_b200.frames import allreduce_gradients, frame_range, gather_frames, shard_frames


def test_frame_range_covers_everything_once():
    for n, world, clip in [(16, 1, 8), (16, 2, 8), (16, 8, 1), (36, 8, 1), (24, 5, 8), (7, 3, 1), (2, 4, 1)]:
        seen = []
        for r in range(world):
            lo, hi = frame_range(n, world, r, clip)
            assert lo % clip == 0 and hi % clip == 0 and lo <= hi
            seen += list(range(lo, hi))
        assert seen == list(range(n))
    with pytest.raises(ValueError):
        frame_range(10, 2, 0, clip_len=4)


def _worker(rank, world, port, n_frames):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_frames * 3 * 2, dtype=torch.float32).view(n_frames, 3, 2)
        (mine,) = shard_frames([full], world, rank)
        out = mine * 2 + 1                                   # stand-in for the per-frame op
        back = gather_frames(out, n_frames)
        assert torch.equal(back, full * 2 + 1)

        lin = torch.nn.Linear(4, 3)
        with torch.no_grad():
            for p in lin.parameters():
                p.fill_(0.5)
        x = torch.full((2, 4), float(rank + 1))
        lin(x).sum().backward()
        n_buckets = allreduce_gradients(lin.parameters(), bucket_bytes=16)   # force several buckets
        assert n_buckets >= 2
        mean_in = sum(range(1, world + 1)) / world
        assert torch.allclose(lin.weight.grad, torch.full((3, 4), 2 * mean_in))
        assert torch.allclose(lin.bias.grad, torch.full((3,), 2.0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [16, 5])
def test_gather_and_allreduce_world2(n_frames):
    port = 29500 + (os.getpid() % 400) + n_frames
    mp.spawn(_worker, args=(2, port, n_frames), nprocs=2, join=True)
