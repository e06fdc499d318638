"""Device-resident replacements for the host sections of the MonoDETR training step.

Each function states the reference lines it stands in for and reproduces their results; `install` attaches them to
live reference objects (monkeypatching instances / module attributes -- the reference tree itself is not edited).
"""
from __future__ import annotations

import math
import sys
import types

import torch

from .lsa import group_lsa

_INDEX_CACHE: dict = {}


def _box_image_index(num_gt_per_img, device):
    """image index of every ground-truth box (boxes are concatenated image after image); cached per size tuple so
    that a steady-state step performs no host-to-device copy for it"""
    key = (tuple(int(n) for n in num_gt_per_img), str(device))
    idx = _INDEX_CACHE.get(key)
    if idx is None:
        if len(_INDEX_CACHE) > 256:
            _INDEX_CACHE.clear()
        flat = [b for b, n in enumerate(key[0]) for _ in range(n)]
        idx = _INDEX_CACHE[key] = torch.tensor(flat, dtype=torch.long).to(device)
    return idx


def _slice_mask(lo, hi, size, device):
    """(nb, size) mask of `range(size)[lo:hi]` for per-box tensors lo / hi with Python's slice rules (a negative
    bound counts from the end, everything is clipped) -- what `t[b, v1:v2, u1:u2] = x` does in the reference loops"""
    def norm(s):
        return torch.where(s < 0, (s + size).clamp_(min=0), s.clamp(max=size))
    pos = torch.arange(size, device=device)[None, :]
    return (pos >= norm(lo)[:, None]) & (pos < norm(hi)[:, None])


def _snap_boxes_(gt_boxes2d):
    """ddn_loss.py:46-48 / balancer.py:66-69: corners snapped outwards IN PLACE (the reference mutates the tensor it was
    given and the second call sees the snapped values), then cast to long"""
    gt_boxes2d[:, :2] = torch.floor(gt_boxes2d[:, :2])
    gt_boxes2d[:, 2:] = torch.ceil(gt_boxes2d[:, 2:])
    return gt_boxes2d.long()


def paint_depth_targets(self, depth_logits, gt_boxes2d, gt_center_depth, num_gt_per_img):
    """DDNLoss.build_target_depth_from_3dcenter (depth_predictor/ddn_loss/ddn_loss.py:42-64) without the per-box
    Python loop: the reference paints the boxes of an image far-to-near, so a pixel ends up with the depth of the
    NEAREST box covering it -- a masked minimum over boxes, then 0 where no box covers."""
    B, _, H, W = depth_logits.shape
    dev = depth_logits.device
    boxes = _snap_boxes_(gt_boxes2d)
    if boxes.shape[0] == 0:
        return torch.zeros((B, H, W), device=dev, dtype=depth_logits.dtype)
    img = _box_image_index(num_gt_per_img, dev)
    inside = _slice_mask(boxes[:, 1], boxes[:, 3], H, dev)[:, :, None] & _slice_mask(boxes[:, 0], boxes[:, 2], W, dev)[:, None, :]
    depth = gt_center_depth.to(depth_logits.dtype)
    cand = torch.where(inside, depth[:, None, None], torch.full((), float("inf"), device=dev, dtype=depth.dtype))
    nearest = torch.full((B, H, W), float("inf"), device=dev, dtype=depth.dtype)
    nearest.scatter_reduce_(0, img[:, None, None].expand_as(cand), cand, reduce="amin")
    return torch.where(torch.isinf(nearest), torch.zeros((), device=dev, dtype=depth.dtype), nearest)


def paint_foreground(gt_boxes2d, shape, num_gt_per_img, downsample_factor=1, device=torch.device("cpu")):
    """compute_fg_mask (depth_predictor/ddn_loss/balancer.py:52-81) without the per-box Python loop."""
    gt_boxes2d /= downsample_factor
    boxes = _snap_boxes_(gt_boxes2d)
    B, H, W = shape
    if boxes.shape[0] == 0:
        return torch.zeros(shape, dtype=torch.bool, device=device)
    img = _box_image_index(num_gt_per_img, device)
    inside = _slice_mask(boxes[:, 1], boxes[:, 3], H, device)[:, :, None] & _slice_mask(boxes[:, 0], boxes[:, 2], W, device)[:, None, :]
    return torch.zeros(shape, dtype=torch.int32, device=device).index_add_(0, img, inside.to(torch.int32)) > 0


def pairwise_l1(a, b):
    """torch.cdist(a, b, p=1) for 2- or 4-column inputs, bit for bit, as three element-wise kernels.  torch's cdist
    kernel takes 1.4 ms for the matcher's 8800 x 128 problem (8.5 ms of GPU time per training step over six calls); it
    reduces over the feature dimension with a shuffle-down tree, i.e. (|d0| + |d2|) + (|d1| + |d3|) for four columns
    -- the order reproduced here (tests/test_step_host.py compares with torch.cdist on the device)."""
    d = (a[:, None, :] - b[None, :, :]).abs()
    if d.shape[-1] == 2:
        return d[..., 0] + d[..., 1]
    if d.shape[-1] == 4:
        return (d[..., 0] + d[..., 2]) + (d[..., 1] + d[..., 3])
    return torch.cdist(a, b, p=1)


class DeviceMatcher:
    """HungarianMatcher.forward (matcher.py:35-104) with the assignment solved on the device.

    The cost matrix is built with the reference's own arithmetic (same operations in the same order, so the same
    fp32 values); instead of `C.cpu()` + scipy per image and group, one kernel launch solves all B x groups
    sub-problems (libmonodetr_step_b200.so, one warp each) and the index tensors stay on the device.  Returns the
    reference's format: a list of (query_index, target_index) int64 tensors per image."""

    def __init__(self, matcher):
        self.matcher = matcher
        self.reference_forward = matcher.forward                 # bound method of the reference module
        mod = sys.modules[type(matcher).__module__]
        self.giou = mod.generalized_box_iou
        self.to_xyxy = mod.box_cxcylrtb_to_xyxy

    @torch.no_grad()
    def __call__(self, outputs, targets, group_num=11):
        boxes = outputs["pred_boxes"]
        if not boxes.is_cuda:
            return self.reference_forward(outputs, targets, group_num=group_num)
        m = self.matcher
        bs, nq = boxes.shape[:2]
        prob = outputs["pred_logits"].flatten(0, 1).sigmoid()
        labels = torch.cat([t["labels"] for t in targets]).long()
        gt = torch.cat([t["boxes_3d"] for t in targets])
        alpha, gamma = 0.25, 2.0                                  # matcher.py:62-66 (focal matching cost)
        neg = (1 - alpha) * (prob ** gamma) * (-(1 - prob + 1e-8).log())
        pos = alpha * ((1 - prob) ** gamma) * (-(prob + 1e-8).log())
        c_class = pos[:, labels] - neg[:, labels]
        flat = boxes.flatten(0, 1)
        c_center = pairwise_l1(flat[:, 0:2], gt[:, 0:2])          # matcher.py:68-72 (torch.cdist, p=1)
        c_bbox = pairwise_l1(flat[:, 2:6], gt[:, 2:6])            # matcher.py:74-78
        c_giou = -self.giou(self.to_xyxy(flat), self.to_xyxy(gt))  # matcher.py:80-83
        cost = m.cost_bbox * c_bbox + m.cost_3dcenter * c_center + m.cost_class * c_class + m.cost_giou * c_giou
        sizes = [len(t["boxes"]) for t in targets]
        pairs = group_lsa(cost.view(bs, nq, -1).float(), sizes, group_num)
        if pairs is None:                                        # sub-problem too large for the kernel: host path
            return self.reference_forward(outputs, targets, group_num=group_num)
        return pairs


def foreach_adamw_step(self, closure=None):
    """AdamW.step of lib/helpers/optimizer_helper.py:68-129 with the per-parameter loop replaced by multi-tensor
    (`torch._foreach_*`) calls: identical update rule, state keys and bias correction -- 9 launches per parameter
    GROUP instead of 9 per parameter (~2700 launches and ~65 ms of host time per step in the reference)."""
    loss = closure() if closure is not None else None
    with torch.no_grad():
        for group in self.param_groups:
            if group["amsgrad"]:
                raise RuntimeError("foreach_adamw_step: amsgrad groups are not handled (use the reference step)")
            beta1, beta2 = group["betas"]
            by_step: dict = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p.data)
                    state["exp_avg_sq"] = torch.zeros_like(p.data)
                state["step"] += 1
                by_step.setdefault(state["step"], []).append(p)
            for step, params in by_step.items():
                grads = [p.grad for p in params]
                exp_avgs = [self.state[p]["exp_avg"] for p in params]
                exp_avg_sqs = [self.state[p]["exp_avg_sq"] for p in params]
                torch._foreach_mul_(exp_avgs, beta1)
                torch._foreach_add_(exp_avgs, grads, alpha=1 - beta1)
                torch._foreach_mul_(exp_avg_sqs, beta2)
                torch._foreach_addcmul_(exp_avg_sqs, grads, grads, value=1 - beta2)
                denoms = torch._foreach_sqrt(exp_avg_sqs)
                torch._foreach_add_(denoms, group["eps"])
                step_size = group["lr"] * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
                update = torch._foreach_mul(params, group["weight_decay"])
                torch._foreach_addcdiv_(update, exp_avgs, denoms, value=1)
                torch._foreach_add_(params, update, alpha=-step_size)
    return loss


def memoized_src_permutation(criterion):
    """SetCriterion._get_src_permutation_idx (monodetr.py:1159-1163) is called by every loss term of every decoder layer
    (27 times per step) and rebuilds the same two concatenations from 16 + 16 small tensors each time (~900 tiny
    launches per step).  One result per matching: for a DeviceMatcher result the flat tensors already exist."""
    original = criterion._get_src_permutation_idx
    cache: dict = {}

    def lookup(indices):
        hit = cache.get(id(indices))
        if hit is None or hit[0] is not indices:
            if len(cache) > 8:
                cache.clear()
            flat = getattr(indices, "query_flat", None)
            if flat is not None:
                dev = flat.device
                batch = torch.repeat_interleave(_arange(len(indices), dev), _counts(indices.per_image, dev),
                                                output_size=flat.numel())
                value = (batch, flat)
            else:
                value = original(indices)
            hit = cache[id(indices)] = (indices, value)
        return hit[1]
    return lookup


def _arange(n, device):
    key = ("arange", n, str(device))
    t = _INDEX_CACHE.get(key)
    if t is None:
        t = _INDEX_CACHE[key] = torch.arange(n, device=device)
    return t


def _counts(per_image, device):
    key = ("counts", tuple(per_image), str(device))
    t = _INDEX_CACHE.get(key)
    if t is None:
        if len(_INDEX_CACHE) > 256:
            _INDEX_CACHE.clear()
        t = _INDEX_CACHE[key] = torch.tensor(list(per_image), dtype=torch.long).to(device)
    return t


def fuse_dense_attention(model):
    """the decoder's two dense attentions -- depth cross-attention (550 queries x 1920 depth
    positions) and the grouped self-attention (reference depthaware_transformer.py:455-503) -- are
    ``nn.MultiheadAttention`` modules called as ``mha(q, k, v, ...)[0]``: the averaged attention weights they return
    by default are thrown away, but asking for them forces PyTorch onto the path that materialises the
    (N*heads, Lq, Lk) probability matrix (540 MB in fp32 for the depth attention at batch 16) in forward and backward.
    With ``need_weights=False`` the same module computes the same output through the fused
    ``scaled_dot_product_attention`` kernels (library code: it stays cuDNN / PyTorch's, this only selects it).
    Returns the number of modules patched.  Outputs agree to fp32 rounding (tests/test_step_host.py); with dropout
    active both paths drop attention probabilities with p = module.dropout, from different random streams.

    MEASURED on B200 (profiles/r02_training_step.md): in fp32 -- the reference's precision -- this is a LOSS.  PyTorch
    2.11 has no flash kernel for fp32; SDPA falls to the sm80 SIMT memory-efficient kernels
    (fmha_cutlassF/B_f32_aligned_64x64), 7.1 + 17.0 ms per step against ~5 ms for the materialising path, and the step
    goes from 158.7 to 161.7 ms.  It pays only under bf16 autocast.  Hence opt-in (install(..., attention=True),
    --host-opt all,sdpa), not part of the default set."""
    n = 0
    for layer in model.modules():
        for name in ("cross_attn_depth", "self_attn"):
            mha = getattr(layer, name, None)
            if isinstance(mha, torch.nn.MultiheadAttention) and not getattr(mha, "_msda_fused_sdpa", False):
                orig = mha.forward

                def fwd(query, key, value, *args, _orig=orig, **kw):
                    if len(args) >= 2:                            # need_weights passed positionally: leave the call alone
                        return _orig(query, key, value, *args, **kw)
                    kw["need_weights"] = False
                    return _orig(query, key, value, *args, **kw)

                mha.forward = fwd
                mha._msda_fused_sdpa = True
                n += 1
    return n


def install(criterion=None, optimizer=None, matcher=True, ddn=True, adamw=True, model=None, attention=False, frozen_bn=True):
    """Attach the device-resident sections to live reference objects; returns the names of what was installed."""
    done = []
    if model is not None and frozen_bn:
        from .frozen_bn import fuse_frozen_bn
        if fuse_frozen_bn(model):
            done.append("frozen_bn")
    if model is not None and attention and fuse_dense_attention(model):
        done.append("sdpa")
    if criterion is not None and matcher and hasattr(criterion, "matcher"):
        dm = DeviceMatcher(criterion.matcher)
        criterion.matcher.forward = dm                            # nn.Module.__call__ dispatches to the instance attribute
        if hasattr(criterion, "_get_src_permutation_idx"):
            criterion._get_src_permutation_idx = memoized_src_permutation(criterion)
        done.append("matcher")
    if criterion is not None and ddn and hasattr(criterion, "ddn_loss"):
        criterion.ddn_loss.build_target_depth_from_3dcenter = types.MethodType(paint_depth_targets, criterion.ddn_loss)
        balancer_mod = sys.modules[type(criterion.ddn_loss.balancer).__module__]
        balancer_mod.compute_fg_mask = paint_foreground
        done.append("ddn_loss")
    if optimizer is not None and adamw and type(optimizer).__name__ == "AdamW" and "amsgrad" in optimizer.defaults \
            and not any(g["amsgrad"] for g in optimizer.param_groups) and type(optimizer).__module__ != "torch.optim.adamw":
        optimizer.step = types.MethodType(foreach_adamw_step, optimizer)
        done.append("adamw")
    return done
