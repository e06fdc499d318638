"""monosowa_b200 -- B200-native (sm_100a) MultiScaleDeformableAttention behind the MonoDETR API.

Importing the package loads libmsda_b200.so (building it with nvcc if absent) and registers the
``msda::forward`` / ``msda::backward`` torch.library ops.  There is no CPU or PyTorch fallback.
"""
from . import _lib  # noqa: F401  (fails loudly if the CUDA library is missing)
from .ops import MSDeformAttn, MSDeformAttn_cross, MSDeformAttnFunction, MultiheadAttention  # noqa: F401
from .host import host_step  # noqa: F401

__all__ = ["MSDeformAttnFunction", "MSDeformAttn", "MSDeformAttn_cross", "MultiheadAttention", "host_step"]
__version__ = "0.1.0"
