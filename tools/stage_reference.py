"""Stage the reference's own Python sources for the GPU box (which has no /root/reference).

    python tools/stage_reference.py        # -> baseline/_ref/soc/...   (git-ignored, shipped by gpurun)

TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing is edited and nothing staged is imported by the product:
the files are copied byte for byte from /root/reference into the git-ignored ``baseline/_ref/soc/`` together
with a MANIFEST of their sha256 sums, so that on the B200

  * ``tests/test_gpu_reference_stack.py`` can run the UNMODIFIED ``models/deformable_transformer.py`` ->
    ``models/ops/modules/ms_deform_attn.py`` -> ``models/ops/functions/ms_deform_attn_func.py`` stack on this
    repo's kernels (their ``import MultiScaleDeformableAttention as MSDA`` resolves to the shim at the repo root)
    and compare it with the same stack on the reference's own CUDA op (oracle/_ref);
  * ``bench.py``'s CPU arm can time the reference's ``ms_deform_attn_core_pytorch`` itself
    (``cpu_baseline.kind = "reference"``) instead of the restatement in oracle/.

No ``models/__init__.py`` is staged (the reference's imports all of SOC, timm and pycocotools): ``models`` is
used as a namespace package.
"""
from __future__ import annotations

import hashlib
import json
import shutil
from pathlib import Path

REF = Path("/root/reference")
ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "baseline" / "_ref" / "soc"
FILES = [
    "misc.py",
    "models/deformable_transformer.py",
    "models/ops/test.py",
    "models/ops/functions/__init__.py",
    "models/ops/functions/ms_deform_attn_func.py",
    "models/ops/modules/__init__.py",
    "models/ops/modules/ms_deform_attn.py",
    # the rest of the SOC model, for the whole-model training step of BASELINE.json configs[2] (tools/soc_step.py)
    "utils.py",
    "configs/a2d_sentences.yaml",
    "models/soc.py",
    "models/backbone.py",
    "models/video_swin_transformer.py",
    "models/position_encoding.py",
    "models/vla.py",
    "models/voc.py",
    "models/segmentation.py",
    "models/criterion.py",
    "models/matcher.py",
    "models/postprocessing.py",
]


def stage(force: bool = False) -> Path | None:
    if not REF.exists():
        return OUT if (OUT / "MANIFEST.json").exists() else None
    manifest = {}
    for rel in FILES:
        src, dst = REF / rel, OUT / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        if force or not dst.exists() or dst.read_bytes() != src.read_bytes():
            shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(dst.read_bytes()).hexdigest()
    (OUT / "MANIFEST.json").write_text(json.dumps({"source": str(REF), "sha256": manifest}, indent=1))
    return OUT


def staged_root() -> Path | None:
    """The staged tree, or None when it was never staged (fresh clone without /root/reference)."""
    return OUT if (OUT / "MANIFEST.json").exists() else None


if __name__ == "__main__":
    print(stage(force=True))
