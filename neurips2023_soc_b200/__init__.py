"""B200-native (sm_100a) multi-scale deformable attention for RobertLuo1/NeurIPS2023_SOC.

One hot path, behind the reference's own operator boundary:

* ``MultiScaleDeformableAttention`` (repo root) / ``msda_ext`` -- ``ms_deform_attn_forward`` and
  ``ms_deform_attn_backward`` with the reference extension's signatures;
* ``functions.MSDeformAttnFunction`` -- the autograd function;
* ``modules.MSDeformAttn`` -- the ``nn.Module`` (same parameters and state-dict keys);
* ``modules.DeformableTransformerEncoderLayer`` / ``DeformableTransformerEncoder`` -- the op's caller in the encoder
  with the elementwise glue fused around it, same constructor, parameters and forward signatures;
* ``frames`` -- frame sharding of a batch over the ranks of one node.

The kernels live in ``csrc/`` and are reached through the C ABI in ``include/msda_b200.h``
(``libmsda_b200.so``).  There is no CPU or PyTorch fallback: without the library, or on CPU
tensors, calls raise.
"""
from .functions import MSDeformAttnFunction, MSDeformAttnFusedFunction
from .modules import DeformableTransformerEncoder, DeformableTransformerEncoderLayer, MSDeformAttn

__all__ = ["MSDeformAttnFunction", "MSDeformAttnFusedFunction", "MSDeformAttn", "DeformableTransformerEncoderLayer",
           "DeformableTransformerEncoder"]
__version__ = "0.1.0"
