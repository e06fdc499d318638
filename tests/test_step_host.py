"""SURVEY.md 8 row f3 -- device-resident replacements of the training step's host sections
(monosowa_b200/step_host).  CPU tests check the vectorised painters and the multi-tensor AdamW against literal
restatements of the reference loops; GPU tests check the assignment kernel against scipy (the reference's solver)
and the whole matcher against the staged reference matcher when baseline/_ref is present."""
import math
import os
import sys

import pytest
import torch

from conftest import ROOT


# ---- literal restatements of the reference loops (test-side checkers) -----------------------------------------
def ref_depth_targets(depth_logits, gt_boxes2d, gt_center_depth, num_gt_per_img):
    """depth_predictor/ddn_loss/ddn_loss.py:42-64, statement by statement"""
    B, _, H, W = depth_logits.shape
    depth_maps = torch.zeros((B, H, W), dtype=depth_logits.dtype)
    gt_boxes2d[:, :2] = torch.floor(gt_boxes2d[:, :2])
    gt_boxes2d[:, 2:] = torch.ceil(gt_boxes2d[:, 2:])
    boxes = gt_boxes2d.long().split(num_gt_per_img, dim=0)
    depths = gt_center_depth.split(num_gt_per_img, dim=0)
    for b in range(len(boxes)):
        d, order = torch.sort(depths[b], dim=0, descending=True)
        bb = boxes[b][order]
        for n in range(bb.shape[0]):
            u1, v1, u2, v2 = (int(x) for x in bb[n])
            depth_maps[b, v1:v2, u1:u2] = d[n]
    return depth_maps


def ref_fg_mask(gt_boxes2d, shape, num_gt_per_img, downsample_factor=1):
    """depth_predictor/ddn_loss/balancer.py:52-81"""
    fg = torch.zeros(shape, dtype=torch.bool)
    gt_boxes2d /= downsample_factor
    gt_boxes2d[:, :2] = torch.floor(gt_boxes2d[:, :2])
    gt_boxes2d[:, 2:] = torch.ceil(gt_boxes2d[:, 2:])
    boxes = gt_boxes2d.long().split(num_gt_per_img, dim=0)
    for b in range(len(boxes)):
        for n in range(boxes[b].shape[0]):
            u1, v1, u2, v2 = (int(x) for x in boxes[b][n])
            fg[b, v1:v2, u1:u2] = True
    return fg


def _boxes(seed, num_gt, H=24, W=80, wild=False):
    g = torch.Generator().manual_seed(seed)
    n = sum(num_gt)
    lo = torch.rand(n, 2, generator=g) * torch.tensor([W * 0.8, H * 0.8])
    wh = torch.rand(n, 2, generator=g) * torch.tensor([W * 0.3, H * 0.5])
    boxes = torch.cat([lo, lo + wh], 1)
    if wild:                                                # boxes hanging over every border, degenerate and inverted ones
        boxes = boxes + (torch.rand(n, 4, generator=g) - 0.5) * torch.tensor([W, H, W, H]) * 1.5
    depth = torch.rand(n, generator=g) * 60
    if n > 3:
        depth[1] = depth[0]                                 # equal depths: the painters must still agree
    return boxes, depth


@pytest.mark.parametrize("num_gt,wild", [([8] * 4, False), ([3, 0, 5, 1], False), ([6, 7, 2], True), ([0, 0], False), ([12], True)])
def test_painters_match_reference_loops(num_gt, wild):
    from monosowa_b200.step_host import paint_depth_targets, paint_foreground
    B, H, W = len(num_gt), 24, 80
    boxes, depth = _boxes(11 + len(num_gt), num_gt, H, W, wild)
    logits = torch.zeros(B, 81, H, W)
    b1, b2 = boxes.clone(), boxes.clone()
    want = ref_depth_targets(logits, b1, depth, num_gt)
    got = paint_depth_targets(None, logits, b2, depth, num_gt)
    assert torch.equal(got, want)
    assert torch.equal(b1, b2)                              # same in-place snapping of the caller's tensor
    want_fg = ref_fg_mask(b1, (B, H, W), num_gt)
    got_fg = paint_foreground(b2, (B, H, W), num_gt, device=torch.device("cpu"))
    assert torch.equal(got_fg, want_fg)


def test_foreach_adamw_matches_reference_loop():
    """optimizer_helper.py:76-127 restated per parameter vs the multi-tensor step, incl. a parameter that
    receives no gradient in some steps (its step counter lags) and the two weight-decay groups"""
    from monosowa_b200.step_host import foreach_adamw_step
    import types

    class AdamW(torch.optim.Optimizer):                     # same state layout / defaults as the reference class
        def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
            super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))

        def step(self):
            for group in self.param_groups:
                for p in group["params"]:
                    if p.grad is None:
                        continue
                    grad, state = p.grad.data, self.state[p]
                    if len(state) == 0:
                        state["step"] = 0
                        state["exp_avg"] = torch.zeros_like(p.data)
                        state["exp_avg_sq"] = torch.zeros_like(p.data)
                    beta1, beta2 = group["betas"]
                    state["step"] += 1
                    state["exp_avg"].mul_(beta1).add_(grad, alpha=1 - beta1)
                    state["exp_avg_sq"].mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
                    denom = state["exp_avg_sq"].sqrt().add_(group["eps"])
                    step_size = group["lr"] * math.sqrt(1 - beta2 ** state["step"]) / (1 - beta1 ** state["step"])
                    p.data.add_(torch.mul(p.data, group["weight_decay"]).addcdiv_(state["exp_avg"], denom, value=1), alpha=-step_size)

    def make():
        torch.manual_seed(0)
        ps = [torch.nn.Parameter(torch.randn(s)) for s in ((7, 5), (5,), (3, 3, 2), (1,))]
        return ps, AdamW([{"params": ps[1::2], "weight_decay": 0}, {"params": ps[0::2], "weight_decay": 1e-4}], lr=2e-4)

    pa, oa = make()
    pb, ob = make()
    ob.step = types.MethodType(foreach_adamw_step, ob)
    g = torch.Generator().manual_seed(1)
    for it in range(6):
        for i, (a, b) in enumerate(zip(pa, pb)):
            if i == 2 and it % 2 == 1:
                a.grad = b.grad = None                      # this parameter skips every other step
            else:
                gr = torch.randn(a.shape, generator=g)
                a.grad, b.grad = gr.clone(), gr.clone()
        oa.step(); ob.step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-8)
        assert oa.state[a]["step"] == ob.state[b]["step"]
        assert torch.allclose(oa.state[a]["exp_avg_sq"], ob.state[b]["exp_avg_sq"], rtol=1e-6, atol=0)


# names DeviceMatcher looks up in the matcher's defining module (the reference imports them from utils.box_ops)
def generalized_box_iou(a, b):
    raise AssertionError("not reached on CPU tensors")


box_cxcylrtb_to_xyxy = generalized_box_iou


def test_install_wires_live_objects_and_keeps_the_host_path_for_cpu_tensors():
    """install() patches instances only; on CPU tensors the matcher keeps the reference's own host path (there is no
    CPU implementation of the device kernel), and the AdamW replacement is attached only to the reference's class shape"""
    import types
    from monosowa_b200 import step_host

    class Matcher(torch.nn.Module):
        cost_class = cost_bbox = cost_3dcenter = cost_giou = 1.0

        def forward(self, outputs, targets, group_num=11):
            return "reference-host-path"

    class Balancer(torch.nn.Module):
        pass

    class DDN(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.balancer = Balancer()

        def build_target_depth_from_3dcenter(self, *a):
            return "loop"

    class Criterion(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.matcher, self.ddn_loss = Matcher(), DDN()

        def _get_src_permutation_idx(self, indices):
            return "original"

    class AdamW(torch.optim.Optimizer):
        def __init__(self, params):
            super().__init__(params, dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False))

        def step(self):
            return "loop"

    crit, opt = Criterion(), AdamW([torch.nn.Parameter(torch.zeros(2))])
    done = step_host.install(crit, opt)
    assert done == ["matcher", "ddn_loss", "adamw"]
    outputs = {"pred_boxes": torch.zeros(1, 11, 6), "pred_logits": torch.zeros(1, 11, 3)}
    assert crit.matcher(outputs, [], group_num=11) == "reference-host-path"
    assert crit.ddn_loss.build_target_depth_from_3dcenter.__func__ is step_host.paint_depth_targets
    assert sys.modules[Balancer.__module__].compute_fg_mask is step_host.paint_foreground
    assert opt.step.__func__ is step_host.foreach_adamw_step
    assert crit._get_src_permutation_idx([(torch.tensor([1]), torch.tensor([0]))]) == "original"   # plain lists: original code
    assert step_host.install(crit, torch.optim.AdamW([torch.nn.Parameter(torch.zeros(2))]), matcher=False, ddn=False) == []


def test_host_step_rejects_what_it_cannot_run():
    from monosowa_b200 import host_step
    v = torch.zeros(1, 4, 2, 32); loc = torch.zeros(1, 3, 2, 1, 4, 2); at = torch.zeros(1, 3, 2, 1, 4); g = torch.zeros(1, 3, 64)
    sh, lsi = torch.tensor([[2, 2]]), torch.tensor([0])
    with pytest.raises(NotImplementedError):                # no CUDA tensor anywhere: there is no CPU implementation
        host_step(v, sh, lsi, loc, at, g)


def test_step_library_exports_declared_symbols():
    import ctypes
    import re
    from monosowa_b200.step_host import lsa
    hdr = open(os.path.join(ROOT, "include", "monodetr_step_b200.h")).read()
    declared = set(re.findall(r"\b(detr_[a-z0-9_]+)\s*\(", hdr))
    assert declared == {"detr_group_lsa_f32", "detr_group_lsa_status_f32", "detr_frozen_bn_act_f32",
                        "detr_frozen_bn_act_backward_f32", "detr_step_last_error"}
    raw = ctypes.CDLL(lsa.LIB_PATH) if os.path.exists(lsa.LIB_PATH) else lsa.lib()
    for name in declared:
        assert hasattr(raw, name)


# ---- GPU ----------------------------------------------------------------------------------------------------------
def _scipy_pairs(cost, sizes, groups):
    """matcher.py:87-104 on the host (the reference's own solver)"""
    import numpy as np
    from scipy.optimize import linear_sum_assignment
    C = cost.cpu()
    nq = C.shape[1] // groups
    out = []
    for b, c in enumerate(C.split(sizes, -1)):
        qi, ti = [], []
        for g in range(groups):
            r, col = linear_sum_assignment(c[b, g * nq:(g + 1) * nq])
            qi.append(r + g * nq); ti.append(col)
        out.append((torch.as_tensor(np.concatenate(qi)), torch.as_tensor(np.concatenate(ti))))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("case", [
    (16, 550, 11, [8] * 16),                       # the training shape: 50 queries per group, 8 targets per image
    (3, 50, 1, [8, 0, 5]),                         # eval: one group; an image without targets
    (2, 40, 4, [10, 17]),                          # 10 queries per group: one image square, one with MORE targets than queries
    (1, 64, 2, [50]),                              # more targets than queries per group (rows = queries)
    (4, 96, 3, [1, 32, 31, 33]),                   # around the warp width
    (1, 300, 1, [120]),                            # several columns per lane
], ids=lambda c: f"B{c[0]}Q{c[1]}G{c[2]}")
def test_group_lsa_matches_scipy(cuda_device, case):
    from monosowa_b200.step_host import group_lsa
    B, Q, groups, sizes = case
    g = torch.Generator().manual_seed(B * 1000 + Q)
    cost = (torch.randn(B, Q, sum(sizes), generator=g) * 3).to(cuda_device)
    got = group_lsa(cost, sizes, groups)
    want = _scipy_pairs(cost, sizes, groups)
    assert len(got) == B
    for (gq, gt), (wq, wt) in zip(got, want):
        assert gq.dtype == torch.int64 and gq.is_cuda
        assert torch.equal(gq.cpu(), wq) and torch.equal(gt.cpu(), wt)


@pytest.mark.gpu
def test_group_lsa_reports_non_finite_costs_like_scipy(cuda_device):
    """scipy.optimize.linear_sum_assignment raises ValueError on NaN / Inf costs (reference matcher.py:101), which stops
    a diverged run; the device matcher raises the same error -- when the flag copied behind the kernel has arrived,
    i.e. at the next call or at an explicit check_status()"""
    from scipy.optimize import linear_sum_assignment
    from monosowa_b200.step_host import lsa
    cost = torch.rand(2, 20, 12, device=cuda_device)
    assert lsa.group_lsa(cost, [5, 7], 2) is not None
    lsa.check_status()                                           # finite costs: nothing to report
    bad = cost.clone()
    bad[1, :, 5:] = float("nan")                                 # every cost of image 1 is NaN
    with pytest.raises(ValueError):
        linear_sum_assignment(bad[1, :10, 5:].cpu())
    lsa.group_lsa(bad, [5, 7], 2)                                # returns (no synchronisation inside the call) ...
    with pytest.raises(ValueError, match="invalid numeric entries"):
        lsa.check_status()                                       # ... and the verdict follows
    lsa.check_status()                                           # reported once
    lsa.group_lsa(bad, [5, 7], 2)
    torch.cuda.synchronize()
    with pytest.raises(ValueError, match="invalid numeric entries"):
        lsa.group_lsa(cost, [5, 7], 2)                           # the next call raises for the previous one


@pytest.mark.gpu
def test_group_lsa_cost_is_optimal_with_ties(cuda_device):
    """integer costs have many exact ties: the assignment may differ from scipy's, its total cost may not"""
    from monosowa_b200.step_host import group_lsa
    sizes = [6, 9, 3]
    cost = torch.randint(0, 4, (3, 60, sum(sizes)), generator=torch.Generator().manual_seed(3)).float().to(cuda_device)
    got = group_lsa(cost, sizes, 5)
    want = _scipy_pairs(cost, sizes, 5)
    off = 0
    for b, ((gq, gt), (wq, wt)) in enumerate(zip(got, want)):
        c = cost[b, :, off:off + sizes[b]].cpu()
        assert len(set(gq.tolist())) == len(gq) and sorted(gt.tolist()) == sorted(wt.tolist())
        assert float(c[gq.cpu(), gt.cpu()].sum()) == float(c[wq, wt].sum())
        off += sizes[b]


@pytest.mark.gpu
@pytest.mark.parametrize("cols", [2, 4])
def test_pairwise_l1_is_bitwise_cdist(cuda_device, cols):
    from monosowa_b200.step_host import pairwise_l1
    g = torch.Generator().manual_seed(cols)
    for n, m in ((8800, 128), (550, 8), (37, 1), (5, 300)):
        a = torch.rand(n, cols, generator=g).to(cuda_device)
        b = torch.rand(m, cols, generator=g).to(cuda_device)
        assert torch.equal(pairwise_l1(a, b), torch.cdist(a, b, p=1)), (n, m, cols)


def _staged_reference():
    ref = os.path.join(ROOT, "baseline", "_ref", "MonoDETR")
    return ref if os.path.isdir(os.path.join(ref, "lib")) else None


@pytest.mark.gpu
@pytest.mark.skipif(_staged_reference() is None, reason="baseline/_ref/MonoDETR not staged (tools/stage_reference.py)")
def test_device_matcher_matches_reference_matcher(cuda_device):
    """the unmodified reference HungarianMatcher (staged copy) vs DeviceMatcher on the same random predictions"""
    import importlib.util
    ref = _staged_reference()
    sys.path.insert(0, ref)                                 # matcher.py imports utils.box_ops from the reference tree
    try:                                                    # (loaded by path: the package __init__ would pull in the whole model)
        spec = importlib.util.spec_from_file_location("_ref_matcher", os.path.join(ref, "lib", "models", "monodetr", "matcher.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["_ref_matcher"] = mod
        spec.loader.exec_module(mod)
        HungarianMatcher = mod.HungarianMatcher
    finally:
        sys.path.remove(ref)
    from monosowa_b200.step_host import DeviceMatcher
    g = torch.Generator().manual_seed(5)
    bs, nq, sizes = 4, 550, [8, 3, 0, 11]
    outputs = {"pred_logits": torch.randn(bs, nq, 3, generator=g).to(cuda_device),
               "pred_boxes": (torch.rand(bs, nq, 6, generator=g) * 0.5 + 0.05).to(cuda_device)}
    targets = [{"labels": torch.randint(0, 3, (n,), generator=g).to(cuda_device),
                "boxes": torch.rand(n, 4, generator=g).to(cuda_device),
                "boxes_3d": (torch.rand(n, 6, generator=g) * 0.5 + 0.05).to(cuda_device)} for n in sizes]
    matcher = HungarianMatcher(cost_class=2, cost_bbox=5, cost_3dcenter=10, cost_giou=2)
    want = matcher(outputs, targets, group_num=11)
    got = DeviceMatcher(matcher)(outputs, targets, group_num=11)
    for (gq, gt), (wq, wt) in zip(got, want):
        assert torch.equal(gq.cpu(), wq) and torch.equal(gt.cpu(), wt)


# ------------------------------------------------------------------------------------------------
# the decoder's dense attentions through fused SDPA (step_host.fuse_dense_attention)
# ------------------------------------------------------------------------------------------------
class _Layer(torch.nn.Module):
    """the two attention members of the reference's DepthAwareDecoderLayer (depthaware_transformer.py:399, 404)"""

    def __init__(self, d_model=256, heads=8, dropout=0.0):
        super().__init__()
        self.cross_attn_depth = torch.nn.MultiheadAttention(d_model, heads, dropout=dropout)
        self.self_attn = torch.nn.MultiheadAttention(d_model, heads, dropout=dropout)
        self.other = torch.nn.MultiheadAttention(d_model, heads, dropout=dropout)      # not one of the two: left alone


def test_fuse_dense_attention_patches_exactly_the_two_decoder_attentions():
    from monosowa_b200 import step_host
    layer = _Layer(32, 4)
    assert step_host.fuse_dense_attention(torch.nn.Sequential(layer)) == 2
    assert step_host.fuse_dense_attention(torch.nn.Sequential(layer)) == 0          # idempotent
    q, kv = torch.randn(5, 2, 32), torch.randn(7, 2, 32)
    out, w = layer.cross_attn_depth(q, kv, kv, key_padding_mask=torch.zeros(2, 7, dtype=torch.bool))
    assert w is None and out.shape == (5, 2, 32)                                      # the reference only uses [0]
    assert layer.other(q, kv, kv)[1] is not None
    ref = torch.nn.MultiheadAttention(32, 4)
    ref.load_state_dict(layer.cross_attn_depth.state_dict())
    assert torch.allclose(out, ref(q, kv, kv)[0], atol=1e-6)


@pytest.mark.gpu
def test_fused_sdpa_matches_the_unmodified_attention_layers(cuda_device):
    """depth cross-attention at the training shape (550 queries x 1920 depth positions, batch 16, key padding mask) and
    the grouped self-attention shape (50 x 50, batch 16 * 11 groups): outputs rel-L2 <= 1e-5, gradients <= 1e-4"""
    from monosowa_b200 import step_host
    dev = cuda_device
    torch.manual_seed(5)
    plain, fused = _Layer().to(dev), _Layer().to(dev)
    fused.load_state_dict(plain.state_dict())
    assert step_host.fuse_dense_attention(fused) == 2
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
    for name, lq, lk, n, masked in (("cross_attn_depth", 550, 1920, 16, True), ("self_attn", 50, 50, 176, False)):
        q = torch.randn(lq, n, 256, device=dev, requires_grad=True)
        k = torch.randn(lk, n, 256, device=dev, requires_grad=True)
        mask = None
        if masked:
            mask = torch.zeros(n, lk, dtype=torch.bool, device=dev)
            mask[:, -80:] = True                                   # a padded bottom strip of the depth map
        g = torch.randn(lq, n, 256, device=dev)
        res = []
        for layer in (plain, fused):
            for t in (q, k):
                t.grad = None
            layer.zero_grad()
            mha = getattr(layer, name)
            out = mha(q, k, k, key_padding_mask=mask)[0]
            out.backward(g)
            res.append((out.detach(), q.grad.clone(), k.grad.clone(), mha.in_proj_weight.grad.clone()))
        assert rel(res[1][0], res[0][0]) < 1e-5, name
        for a, b in zip(res[1][1:], res[0][1:]):
            assert rel(a, b) < 1e-4, name


# ------------------------------------------------------------------------------------------------
# FrozenBatchNorm2d (+ identity) (+ ReLU) in one pass (step_host.fuse_frozen_bn): bitwise identical to the reference
# ------------------------------------------------------------------------------------------------
class FrozenBatchNorm2d(torch.nn.Module):
    """the reference's module, restated for the test (MonoDETR/lib/models/monodetr/backbone.py:28-65)"""

    def __init__(self, n, eps=1e-5):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))
        self.eps = eps

    def forward(self, x):
        w = self.weight.reshape(1, -1, 1, 1)
        b = self.bias.reshape(1, -1, 1, 1)
        rv = self.running_var.reshape(1, -1, 1, 1)
        rm = self.running_mean.reshape(1, -1, 1, 1)
        scale = w * (rv + self.eps).rsqrt()
        bias = b - rm * scale
        return x * scale + bias


def _randomise_bn(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, FrozenBatchNorm2d):
            m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.3)
            m.running_mean.copy_(torch.randn(m.bias.shape, generator=g) * 0.2)
            m.running_var.copy_(torch.rand(m.bias.shape, generator=g) + 0.3)


def test_fuse_frozen_bn_patches_and_leaves_cpu_tensors_to_the_reference_code():
    from torchvision.models.resnet import Bottleneck
    from monosowa_b200 import step_host
    block = Bottleneck(16, 4, norm_layer=FrozenBatchNorm2d)
    net = torch.nn.Sequential(block, FrozenBatchNorm2d(16))
    _randomise_bn(net)
    x = torch.randn(2, 16, 5, 8)
    want = net(x)
    assert step_host.fuse_frozen_bn(net) == 5            # bn1..bn3, the block, the trailing norm
    assert step_host.fuse_frozen_bn(net) == 0            # idempotent
    assert torch.equal(net(x), want)                     # CPU input: the reference's own forward runs


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 64, 24, 80), (1, 16, 5, 7), (3, 32, 6, 20)], ids=lambda s: "x".join(map(str, s)))
def test_fused_frozen_bn_is_bitwise_identical(cuda_device, shape):
    """outputs AND gradients of: a bare FrozenBatchNorm2d; torchvision Bottleneck blocks without and with a downsample
    branch (bn + relu twice, bn + identity + relu) -- every arithmetic step keeps the reference's rounding"""
    import copy
    from torchvision.models.resnet import Bottleneck
    from monosowa_b200 import step_host
    dev = cuda_device
    n, c, h, w = shape
    torch.manual_seed(11)
    down = torch.nn.Sequential(torch.nn.Conv2d(c, c, 1, bias=False), FrozenBatchNorm2d(c))
    ref = torch.nn.ModuleList([FrozenBatchNorm2d(c), Bottleneck(c, c // 4, norm_layer=FrozenBatchNorm2d),
                               Bottleneck(c, c // 4, downsample=down, norm_layer=FrozenBatchNorm2d)]).to(dev)
    _randomise_bn(ref)
    fused = copy.deepcopy(ref)
    assert step_host.fuse_frozen_bn(fused) == 1 + 4 + 5
    def run(net, i, x, g):
        net.zero_grad()
        xi = x.clone().requires_grad_(True)
        y = net[i](xi)
        y.backward(g)
        return y.detach(), xi.grad, [p.grad.clone() for p in net[i].parameters()]

    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()

    def same(a, b, b2, what):
        # cuDNN runs its deterministic algorithms here (flag below), so two runs of the REFERENCE agree bit for bit and
        # so must the fused path; should a cuDNN build still not be repeatable (b vs b2), fp32 rounding noise is allowed
        if torch.equal(b, b2):
            assert torch.equal(a, b), f"{what} differs"
        else:
            assert rel(a, b) < 1e-5, f"{what} differs by {rel(a, b)}; the reference's own run-to-run noise is {rel(b2, b)}"

    prev = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True               # convolution backward without run-dependent summation order
    for i in range(3):
        x = torch.randn(n, c, h, w, device=dev)
        g = torch.randn(n, c, h, w, device=dev)
        r1, r2, f = run(ref, i, x, g), run(ref, i, x, g), run(fused, i, x, g)
        assert torch.equal(r1[0], f[0]), f"module {i}: output differs"          # forward: always bit for bit
        same(f[1], r1[1], r2[1], f"module {i}: input gradient")
        for a, b, b2 in zip(f[2], r1[2], r2[2]):
            same(a, b, b2, f"module {i}: a weight gradient")
        if i == 0:
            assert torch.equal(f[1], r1[1]), "a bare FrozenBatchNorm2d must match bit for bit in both directions"
    torch.backends.cudnn.deterministic = prev
    # buffers changed (a checkpoint was loaded): the cached scale / bias must follow
    _randomise_bn(ref, seed=5)
    fused.load_state_dict(ref.state_dict())
    x = torch.randn(n, c, h, w, device=dev)
    assert torch.equal(ref[0](x), fused[0](x))
