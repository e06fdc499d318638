"""MSDeformAttn module -- drop-in for the reference's
MonoDETR/lib/models/monodetr/ops/modules/ms_deform_attn.py:63-162 (class MSDeformAttn).

Same constructor signature, attribute names, parameter names / shapes / initialisation (so
reference checkpoints load with ``load_state_dict``) and the same forward arithmetic:
value_proj -> padding mask -> (N,S,M,D) view; sampling_offsets / attention_weights Linears;
softmax over L*P; sampling locations from 2-dim or 6-dim reference points (:149-155); the MSDA
op; output_proj.  The four Linears stay on cuBLAS; only the gather/scatter op is ours.

``MSDeformAttn_cross`` (reference :164-256) is byte-for-byte the ``conditional=True`` flavour of
the same module and ``MultiheadAttention`` (reference :259-589) is a vendored copy of an old
torch.nn.MultiheadAttention; nothing in the reference instantiates either
(depthaware_transformer.py:11 only imports them), so they are provided as thin aliases.
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn.functional as F
from torch import nn

from ..functions import MSDeformAttnFunction, MSDeformAttnFusedFunction, fused_supported


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


def sampling_locations_from_reference(reference_points, sampling_offsets, spatial_shapes, n_points):
    """reference_points (N,Lq,L,2|6), sampling_offsets (N,Lq,M,L,P,2) -> locations (N,Lq,M,L,P,2).
    Reference arithmetic: ms_deform_attn.py:149-155."""
    last = reference_points.shape[-1]
    ref = reference_points[:, :, None, :, None, :]
    if last == 2:
        wh = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)      # (L,2) as (W,H)
        return ref + sampling_offsets / wh[None, None, None, :, None, :]
    if last == 6:
        extent = ref[..., 2::2] + ref[..., 3::2]                                    # (l+r, t+b)
        return ref[..., :2] + sampling_offsets / n_points * extent * 0.5
    raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(last))


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4, conditional=False):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        if not _is_power_of_2(d_model // n_heads):
            warnings.warn("You'd better set d_model in MSDeformAttn to make the dimension of each attention head "
                          "a power of 2 which is more efficient in our CUDA implementation.")
        self.im2col_step = 64                      # reference :87; kept for API parity (validated, not needed)
        # fold softmax + the reference-point arithmetic into the kernels when the call allows it
        # (all six MSDA calls of a MonoDETR training forward do); set False for the literal path
        self.fuse_preprocessing = True
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.conditional = conditional
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        width = d_model // 2 if conditional else d_model
        self.value_proj = nn.Linear(width, width)
        self.output_proj = nn.Linear(width, width)
        self._reset_parameters()

    def _reset_parameters(self):
        """Reference :106-120: offsets start as a ring of unit directions scaled by point index."""
        with torch.no_grad():
            self.sampling_offsets.weight.zero_()
            theta = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
            ring = torch.stack([theta.cos(), theta.sin()], -1)
            ring = ring / ring.abs().max(-1, keepdim=True)[0]
            ring = ring.view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
            ring = ring * torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, -1, 1)
            self.sampling_offsets.bias = nn.Parameter(ring.reshape(-1))
            self.attention_weights.weight.zero_()
            self.attention_weights.bias.zero_()
            nn.init.xavier_uniform_(self.value_proj.weight)
            self.value_proj.bias.zero_()
            nn.init.xavier_uniform_(self.output_proj.weight)
            self.output_proj.bias.zero_()

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        """query (N,Lq,C); reference_points (N,Lq,L,2) in [0,1] or (N,Lq,L,6); input_flatten (N,S,C);
        input_spatial_shapes (L,2) [(H,W)]; input_level_start_index (L,); input_padding_mask (N,S) bool.
        Returns (N,Lq,C)."""
        n, len_q, _ = query.shape
        _, len_in, _ = input_flatten.shape
        assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == len_in

        value = self.value_proj(input_flatten)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))
        value = value.view(n, len_in, self.n_heads, value.shape[-1] // self.n_heads)
        offsets = self.sampling_offsets(query).view(n, len_q, self.n_heads, self.n_levels, self.n_points, 2)
        weights = self.attention_weights(query).view(n, len_q, self.n_heads, self.n_levels * self.n_points)
        # fused path: 2-dim reference points (with or without a gradient: encoder, decoder layer 0) and 6-dim ones
        # that need no gradient (decoder layers 1+: the reference detaches them, depthaware_transformer.py:613)
        if (self.fuse_preprocessing
                and (reference_points.shape[-1] == 2 or not (reference_points.requires_grad and torch.is_grad_enabled()))
                and fused_supported(value, reference_points, offsets, weights, input_spatial_shapes,
                                    input_level_start_index, self.n_levels, self.n_points)):
            output = MSDeformAttnFusedFunction.apply(value, input_spatial_shapes, input_level_start_index,
                                                     reference_points, offsets, weights)
            return self.output_proj(output)
        if value.dtype == torch.bfloat16:
            # bf16 carries 8 mantissa bits: not enough for sub-pixel coordinates at W=160.  Keep the
            # location / weight arithmetic in fp32 (the bf16 kernels take fp32 loc & weights).
            offsets, weights, reference_points = offsets.float(), weights.float(), reference_points.float()
        weights = F.softmax(weights, -1).view(n, len_q, self.n_heads, self.n_levels, self.n_points)
        locations = sampling_locations_from_reference(reference_points, offsets, input_spatial_shapes, self.n_points)
        output = MSDeformAttnFunction.apply(value, input_spatial_shapes, input_level_start_index,
                                            locations.contiguous(), weights.contiguous(), self.im2col_step)
        return self.output_proj(output)


class MSDeformAttn_cross(MSDeformAttn):
    """Reference :164-256 == MSDeformAttn with half-width value/output projections."""

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__(d_model, n_levels, n_heads, n_points, conditional=True)


MultiheadAttention = nn.MultiheadAttention
