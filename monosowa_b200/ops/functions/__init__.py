from .ms_deform_attn_func import (MSDeformAttnFunction, ms_deform_attn_backward,  # noqa: F401
                                  ms_deform_attn_forward)
