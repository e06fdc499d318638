from .ms_deform_attn import MSDeformAttn, MSDeformAttn_cross, MultiheadAttention  # noqa: F401
