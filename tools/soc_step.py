"""Whole-model SOC training step on synthetic A2D clips (BASELINE.json configs[2]): the reference's UNMODIFIED model
(random-init Video-Swin-T + RoBERTa-base + text fusion + Deformable-DETR encoder/decoder + VOC + mask head, staged
under baseline/_ref/soc by tools/stage_reference.py), its own criterion and matcher, AdamW with the reference's three
parameter groups, gradient clipping -- with this repo's kernels behind `MultiScaleDeformableAttention` -- clips sharded
over the ranks with DistributedDataParallel and the gradient all-reduce over NCCL, as /root/reference/trainer.py:52-54,
138-197 does.  This is the integration harness of SURVEY.md 8f-3; none of it is product code.

What the offline box cannot supply is shimmed, nothing in the reference files is edited:
  * `timm.models.layers` (DropPath, trunc_normal_, to_2tuple) and `pycocotools.mask` -> minimal stand-ins when absent;
  * `RobertaModel.from_pretrained` / `RobertaTokenizerFast.from_pretrained` (no weights or vocabulary offline) ->
    a random-init RoBERTa-base of the same architecture and a tokenizer that hashes words to ids;
  * the YAML config is read with PyYAML instead of ruamel.yaml (main.py:17-20 semantics: {key: {value: ...}}).

    python tools/soc_step.py [--batch 2] [--frames 8] [--height 384 --width 640] [--steps 5] [--amp]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/soc_step.py

Prints one JSON line: ms per step (max over ranks), clips/s, the op's kernels per step, the exposed cost of the
gradient all-reduce (step time with DDP's synchronisation minus the same step under no_sync()) and the bare
all-reduce of the same gradient bytes.
"""
from __future__ import annotations

import argparse
import importlib
import json
import math
import os
import sys
import types
import zlib
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
STAGED = ROOT / "baseline" / "_ref" / "soc"
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


# ---------------------------------------------------------------------------------------- shims
def install_shims() -> None:
    """Stand-ins for what the reference imports but this image does not have (never for anything it has)."""
    import transformers                       # before any stand-in: it probes for optional packages by name
    from transformers import BatchEncoding, RobertaConfig, RobertaModel
    try:
        import timm.models.layers  # noqa: F401
    except Exception:
        from torch import nn

        class DropPath(nn.Module):                      # stochastic depth, as timm's
            def __init__(self, drop_prob: float = 0.0):
                super().__init__()
                self.drop_prob = float(drop_prob)

            def forward(self, x):
                if self.drop_prob == 0.0 or not self.training:
                    return x
                keep = 1.0 - self.drop_prob
                mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
                return x * mask / keep

        def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
            return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

        def to_2tuple(x):
            return tuple(x) if isinstance(x, (tuple, list)) else (x, x)
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.DropPath, layers.trunc_normal_, layers.to_2tuple = DropPath, trunc_normal_, to_2tuple
        timm.models, models.layers = models, layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    try:
        import pycocotools.mask  # noqa: F401
    except Exception:
        coco = types.ModuleType("pycocotools")
        mask = types.ModuleType("pycocotools.mask")      # only used by the evaluation post-processors

        def _unavailable(*a, **k):
            raise RuntimeError("pycocotools is not installed: RLE encoding is unavailable in this harness")
        mask.encode = mask.decode = mask.area = mask.toBbox = _unavailable
        coco.mask = mask
        sys.modules.update({"pycocotools": coco, "pycocotools.mask": mask})

    def random_roberta(cls, *args, **kwargs):            # roberta-base architecture, random weights
        cfg = RobertaConfig(vocab_size=50265, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                            intermediate_size=3072, max_position_embeddings=514, type_vocab_size=1, pad_token_id=1,
                            bos_token_id=0, eos_token_id=2, layer_norm_eps=1e-5)
        return RobertaModel(cfg)

    class HashTokenizer:
        """<s> word ids </s>, padded with 1 to the longest query: the interface SOC.forward_text uses."""

        @classmethod
        def from_pretrained(cls, *args, **kwargs):
            return cls()

        def batch_encode_plus(self, texts, padding="longest", return_tensors="pt"):
            rows = [[0] + [3 + zlib.crc32(w.encode()) % 50000 for w in t.lower().split()] + [2] for t in texts]
            n = max(len(r) for r in rows)
            ids = torch.tensor([r + [1] * (n - len(r)) for r in rows], dtype=torch.long)
            att = torch.tensor([[1] * len(r) + [0] * (n - len(r)) for r in rows], dtype=torch.long)
            return BatchEncoding({"input_ids": ids, "attention_mask": att})

    transformers.RobertaModel.from_pretrained = classmethod(random_roberta)
    transformers.RobertaTokenizerFast.from_pretrained = classmethod(lambda cls, *a, **k: HashTokenizer())


def reference_modules():
    """(models.soc, misc) of the staged reference tree; `models` is a namespace package (its own __init__ is not
    staged), `MultiScaleDeformableAttention` resolves to this repo's shim at the repo root."""
    if not (STAGED / "models" / "soc.py").exists():
        raise SystemExit("reference model not staged: run `python tools/stage_reference.py` where /root/reference exists")
    install_shims()
    if str(STAGED) not in sys.path:
        sys.path.insert(1, str(STAGED))
    soc = importlib.import_module("models.soc")
    misc = importlib.import_module("misc")
    return soc, misc


def load_config(device, **overrides) -> argparse.Namespace:
    import re
    import yaml
    with open(STAGED / "configs" / "a2d_sentences.yaml") as f:
        cfg = {k: v["value"] for k, v in yaml.safe_load(f).items()}           # main.py:18
    # PyYAML (YAML 1.1) reads "5e-5" as a string; ruamel.yaml (1.2), which the reference uses, as a float
    num = re.compile(r"^[+-]?\d+(\.\d*)?[eE][+-]?\d+$")
    cfg = {k: (float(v) if isinstance(v, str) and num.match(v) else v) for k, v in cfg.items()}
    cfg.update(backbone="video-swin-t", backbone_pretrained=False, backbone_pretrained_path=None,
               text_encoder_type="roberta-base", device=device, lr_drop=[15], epochs=40)   # scripts/train_a2d.sh
    cfg.update(overrides)
    return argparse.Namespace(**cfg)


def synthetic_batch(misc, B: int, T: int, H: int, W: int, device, seed: int):
    """One A2D-style batch: clips [T, B, 3, H, W] without padding, one text query per clip, and the single annotated
    (centre) frame per clip with one blob instance (datasets/a2d_sentences/a2d_sentences_dataset.py:200-222), already
    filtered the way trainer.py:157-168 does."""
    g = torch.Generator().manual_seed(seed)
    frames = torch.randn(T, B, 3, H, W, generator=g)
    samples = misc.NestedTensor(frames, torch.zeros(T, B, H, W, dtype=torch.bool)).to(device)
    words = "a person dog car bird walking running left right small large jumping near the".split()
    texts = [" ".join(words[int(i)] for i in torch.randint(0, len(words), (6,), generator=g)) for _ in range(B)]
    targets = []
    for b in range(B):
        cy, cx = (torch.rand(2, generator=g) * 0.5 + 0.25).tolist()
        hh, ww = (torch.rand(2, generator=g) * 0.2 + 0.15).tolist()
        ys = torch.arange(H).view(H, 1) / H
        xs = torch.arange(W).view(1, W) / W
        mask = ((ys - cy).abs() < hh / 2) & ((xs - cx).abs() < ww / 2)
        targets.append({"masks": mask[None].to(device), "boxes": torch.tensor([[cx, cy, ww, hh]], device=device),
                        "size": torch.tensor([H, W], device=device), "orig_size": torch.tensor([H, W], device=device),
                        "is_ref_inst_visible": torch.tensor(True, device=device),
                        "referred_instance_idx": torch.tensor(0, device=device),
                        "labels": torch.zeros(1, dtype=torch.long, device=device)})
    valid_indices = torch.tensor([T // 2 + b * T for b in range(B)], device=device)
    return samples, valid_indices, texts, [tuple(targets)]


def build(device, **overrides):
    soc, misc = reference_modules()
    cfg = load_config(device, **overrides)
    model, criterion, _ = soc.build(cfg)
    return cfg, model.to(device), criterion, misc


def optimizer_for(model, cfg):
    """trainer.py:84-99: backbone and text encoder on their own learning rates."""
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    groups = [
        {"params": [p for n, p in named if "backbone" not in n and "text_encoder" not in n]},
        {"params": [p for n, p in named if "backbone" in n], "lr": cfg.lr_backbone},
        {"params": [p for n, p in named if "text_encoder" in n], "lr": cfg.text_encoder_lr},
    ]
    return torch.optim.AdamW([g for g in groups if g["params"]], lr=cfg.lr, weight_decay=cfg.weight_decay)


def train_step(net, criterion, opt, batch, cfg, amp: bool, sync=None):
    """trainer.py:175-197 (GradScaler is disabled in every shipped config)."""
    samples, valid_indices, texts, targets = batch
    samples = type(samples)(samples.tensors.clone(), samples.mask.clone())      # the model rearranges them in place
    ctx = sync() if sync is not None else _Null()
    with ctx:
        with torch.autocast(net_device(net).type, dtype=torch.bfloat16, enabled=amp):
            outputs = net(samples, valid_indices, texts, targets)
            loss_dict = criterion(outputs, targets)
            wd = criterion.weight_dict
            loss = sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)
        opt.zero_grad(set_to_none=True)
        loss.backward()
    if cfg.clip_max_norm > 0:
        torch.nn.utils.clip_grad_norm_(net.parameters(), cfg.clip_max_norm, error_if_nonfinite=False)
    opt.step()
    return loss.detach()


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def net_device(net):
    return next(net.parameters()).device


def main():
    import torch.distributed as dist
    from torch import nn
    from neurips2023_soc_b200 import _lib
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2, help="clips per GPU (scripts/train_a2d.sh: -bs 2)")
    ap.add_argument("--frames", type=int, default=8, help="frames per clip (-ws 8)")
    ap.add_argument("--height", type=int, default=384)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--amp", action="store_true", help="bf16 autocast (the reference's configs train in fp32)")
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    if not torch.cuda.is_available():
        raise SystemExit("tools/soc_step.py needs a CUDA device: the op has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(42 + rank)                                                 # trainer.py:44-47
    cfg, model, criterion, misc = build(dev)
    n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
    n_all = sum(p.numel() for p in model.parameters())
    net = nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model   # trainer.py:52-54
    opt = optimizer_for(model, cfg)
    net.train()
    criterion.train()
    batch = synthetic_batch(misc, a.batch, a.frames, a.height, a.width, dev, seed=rank)

    def timed(sync, steps):
        for _ in range(2):
            train_step(net, criterion, opt, batch, cfg, a.amp, sync)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = train_step(net, criterion, opt, batch, cfg, a.amp, sync)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), float(loss)

    ms, loss = timed(None, a.steps)
    nosync_ms = bare_ms = None
    if world > 1:
        nosync_ms, _ = timed(net.no_sync, a.steps)                               # same step, gradients left unreduced
        flat = torch.empty(n_train, dtype=torch.float32, device=dev)
        for _ in range(2):
            dist.all_reduce(flat)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dist.all_reduce(flat)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 5], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        bare_ms = float(t.item())
    _lib.profile_enable(True)
    train_step(net, criterion, opt, batch, cfg, a.amp, None)
    torch.cuda.synchronize()
    recs = _lib.profile_read()
    _lib.profile_enable(False)
    if rank == 0:
        by = {}
        for name, t in recs:
            by[name] = by.get(name, 0.0) + t
        print(json.dumps({
            "what": "SOC training step (unmodified reference model + criterion, random init, synthetic A2D clips), "
                    "this repo's kernels behind MultiScaleDeformableAttention",
            "n_gpus": world, "clips_per_gpu": a.batch, "frames_per_clip": a.frames, "frame": [a.height, a.width],
            "amp_bf16": a.amp, "params_total_M": n_all / 1e6, "params_trainable_M": n_train / 1e6,
            "gradient_bytes_MB": n_train * 4 / 1e6, "ms_per_step": ms, "clips_per_s": world * a.batch / (ms * 1e-3),
            "ms_per_step_no_gradient_sync": nosync_ms,
            "allreduce_exposed_ms": None if nosync_ms is None else ms - nosync_ms,
            "allreduce_bare_ms": bare_ms,
            "allreduce_bare_busbw_GBs": None if bare_ms is None else 2 * (world - 1) / world * n_train * 4 / (bare_ms * 1e-3) / 1e9,
            "msda_kernels_ms_per_step": sum(by.values()), "msda_launches_per_step": len(recs),
            "msda_kernels": {k: round(v, 4) for k, v in sorted(by.items(), key=lambda kv: -kv[1])},
            "loss": loss if math.isfinite(loss) else str(loss)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
