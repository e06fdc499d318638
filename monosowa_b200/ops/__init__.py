"""Mirror of the reference package MonoDETR/lib/models/monodetr/ops (functions/, modules/)."""
from .functions import MSDeformAttnFunction  # noqa: F401
from .modules import MSDeformAttn, MSDeformAttn_cross, MultiheadAttention  # noqa: F401
