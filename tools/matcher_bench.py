"""Matcher micro-benchmark at the training shape (16 images x 550 queries in 11 groups, 8 targets per image):
the staged reference HungarianMatcher (host: C.cpu() + scipy) vs DeviceMatcher vs the assignment kernel alone."""
import importlib.util
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monosowa_b200.step_host import DeviceMatcher, group_lsa  # noqa: E402

ref = os.path.join(ROOT, "baseline", "_ref", "MonoDETR")
sys.path.insert(0, ref)
spec = importlib.util.spec_from_file_location("_ref_matcher", os.path.join(ref, "lib", "models", "monodetr", "matcher.py"))
mod = importlib.util.module_from_spec(spec); sys.modules["_ref_matcher"] = mod; spec.loader.exec_module(mod)

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
bs, nq, sizes = 16, 550, [8] * 16
outputs = {"pred_logits": torch.randn(bs, nq, 3, generator=g).to(dev), "pred_boxes": (torch.rand(bs, nq, 6, generator=g) * 0.5 + 0.05).to(dev)}
targets = [{"labels": torch.randint(0, 3, (n,), generator=g).to(dev), "boxes": torch.rand(n, 4, generator=g).to(dev),
            "boxes_3d": (torch.rand(n, 6, generator=g) * 0.5 + 0.05).to(dev)} for n in sizes]
matcher = mod.HungarianMatcher(cost_class=2, cost_bbox=5, cost_3dcenter=10, cost_giou=2)
dm = DeviceMatcher(matcher)
cost = torch.randn(bs, nq, sum(sizes), device=dev)


def wall(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t_issue = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return round(t_issue, 3), round((time.perf_counter() - t0) / n * 1e3, 3)


def gpu_ms(fn, n=30):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 3)


print(json.dumps({
    "reference_matcher_ms(issue, total)": wall(lambda: matcher(outputs, targets, group_num=11)),
    "device_matcher_ms(issue, total)": wall(lambda: dm(outputs, targets, group_num=11)),
    "group_lsa_call_ms(issue, total)": wall(lambda: group_lsa(cost, sizes, 11)),
    "group_lsa_gpu_ms": gpu_ms(lambda: group_lsa(cost, sizes, 11)),
}))
