from .ms_deform_attn_func import MSDeformAttnFunction, MSDeformAttnFusedFunction

__all__ = ["MSDeformAttnFunction", "MSDeformAttnFusedFunction"]
