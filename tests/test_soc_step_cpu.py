"""Plumbing of the whole-model harness (tools/soc_step.py, BASELINE.json configs[2]) on CPU: the staged, unmodified
reference SOC model is built with the offline shims (random-init Video-Swin-T / RoBERTa, hash tokenizer, PyYAML
config), one synthetic A2D batch goes forward, through the reference's criterion and matcher, backward and through an
AdamW step.  There is no GPU here, so the two extension entry points are replaced by the CPU oracle for the
duration of the test (as in test_dropin_cpu.py); what is being checked is the harness, not the kernels.  Skips when
the reference model was not staged (python tools/stage_reference.py in the build container)."""
import math
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from neurips2023_soc_b200 import msda_ext
from oracle import msda_oracle as O

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.skipif(not (ROOT / "baseline" / "_ref" / "soc" / "models" / "soc.py").exists(),
                                reason="reference model not staged")


def _oracle_forward(value, shapes, lsi, loc, attn, im2col_step, flags=None, want_index=False, **kw):
    out = torch.from_numpy(O.forward_c(value, shapes, lsi, loc, attn, dtype=np.float32))
    return (out, None) if want_index else out


def _oracle_backward(value, shapes, lsi, loc, attn, grad_out, im2col_step, flags=None, index=None, **kw):
    return [torch.from_numpy(a) for a in O.backward_c(value, shapes, lsi, loc, attn, grad_out, dtype=np.float32)]


def test_whole_model_training_step_runs(monkeypatch):
    import MultiScaleDeformableAttention as shim
    for target in (msda_ext, shim):
        monkeypatch.setattr(target, "ms_deform_attn_forward", _oracle_forward)
        monkeypatch.setattr(target, "ms_deform_attn_backward", _oracle_backward)
    saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.") or k in ("misc", "utils")}
    for k in saved:
        del sys.modules[k]
    try:
        from tools import soc_step
        cfg, model, criterion, misc = soc_step.build(torch.device("cpu"))
        assert type(model.transformer.encoder.layers[0].self_attn).__module__ == "models.ops.modules.ms_deform_attn"
        trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
        assert 49e6 < trainable < 50e6                      # 49.48 M trainable parameters (frozen RoBERTa)
        assert cfg.lr == 5e-5 and cfg.DeformTransformer["num_queries"] == 20
        opt = soc_step.optimizer_for(model, cfg)
        model.train()
        criterion.train()
        batch = soc_step.synthetic_batch(misc, 1, 2, 64, 96, torch.device("cpu"), seed=0)
        before = model.transformer.encoder.layers[0].self_attn.value_proj.weight.detach().clone()
        loss = soc_step.train_step(model, criterion, opt, batch, cfg, amp=False)
        assert math.isfinite(float(loss))
        assert not torch.equal(before, model.transformer.encoder.layers[0].self_attn.value_proj.weight)   # the step moved it
    finally:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k in ("misc", "utils")]:
            del sys.modules[k]
        sys.modules.update(saved)
        for p in list(sys.path):
            if p.endswith("baseline/_ref/soc"):
                sys.path.remove(p)
