"""Drop-in for the reference's compiled extension of the same name.

The reference does ``import MultiScaleDeformableAttention as MSDA``
(/root/reference/models/ops/functions/ms_deform_attn_func.py:18) and calls exactly two
functions on it (:25-26, :35-36; exported at /root/reference/models/ops/src/vision.cpp:13-16).
With this module first on ``sys.path`` the reference's own
``MSDeformAttnFunction`` runs on the sm_100a kernels unchanged.
"""
from neurips2023_soc_b200.msda_ext import ms_deform_attn_backward, ms_deform_attn_forward

__all__ = ["ms_deform_attn_forward", "ms_deform_attn_backward"]
