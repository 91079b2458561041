from .encoder_layer import DeformableTransformerEncoder, DeformableTransformerEncoderLayer
from .ms_deform_attn import MSDeformAttn

__all__ = ["MSDeformAttn", "DeformableTransformerEncoderLayer", "DeformableTransformerEncoder"]
