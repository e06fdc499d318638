"""MSDeformAttn MODULE forward+backward (4 Linears on cuBLAS + pre-processing + the op), fused vs literal
pre-processing, encoder shape of BASELINE.json configs[1].
    python tools/module_bench.py [--batch 16] [--iters 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import monosowa_b200 as msda  # noqa: E402
from monosowa_b200 import workloads as W  # noqa: E402
from tools.sweep import timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--set", default="", help="comma-separated key=value for msda_set_tuning")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    for kv in filter(None, a.set.split(",")):
        k_, v_ = kv.split("=")
        msda._lib.set_tuning(k_, int(v_))
    torch.manual_seed(0)
    sh, lsi = W.level_tensors(W.KITTI, dev)
    S = 10200
    mod = msda.MSDeformAttn().to(dev)
    with torch.no_grad():                                   # trained-like: offsets and weights depend on the query
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.05)
    q = torch.randn(a.batch, S, 256, device=dev, requires_grad=True)
    src = torch.randn(a.batch, S, 256, device=dev, requires_grad=True)
    ref = W.encoder_reference_points(W.KITTI, dev)[None].expand(a.batch, -1, -1, -1).contiguous()
    g = torch.randn(a.batch, S, 256, device=dev)
    bench(mod, "encoder self-attention (2-dim pixel-centre references)", q, ref, src, g, sh, lsi, a)
    # the decoder's calls: 550 training queries; layer 0 feeds learned 2-dim references WITH a gradient,
    # layers 1-2 feed 6-dim boxes, detached (reference depthaware_transformer.py:286, 565-613)
    qd = torch.randn(a.batch, 550, 256, device=dev, requires_grad=True)
    gd = torch.randn(a.batch, 550, 256, device=dev)
    ref2 = torch.rand(a.batch, 550, 1, 2, device=dev).expand(-1, -1, 4, -1).contiguous().requires_grad_(True)
    ref6 = torch.cat([ref2.detach(), torch.rand(a.batch, 550, 4, 4, device=dev) * 0.2 + 0.02], -1)
    bench(mod, "decoder layer 0 (550 queries, 2-dim learned references with gradient)", qd, ref2, src, gd, sh, lsi, a)
    bench(mod, "decoder layers 1-2 (550 queries, 6-dim detached boxes)", qd, ref6, src, gd, sh, lsi, a)


def bench(mod, call, q, ref, src, g, sh, lsi, a):
    S = q.shape[1]
    for amp in (False, True):
        for fused in (False, True):
            mod.fuse_preprocessing = fused

            def fwd():
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                    return mod(q, ref, src, sh, lsi)

            def fwd_bwd():
                out = fwd()
                out.backward(g.to(out.dtype))
                q.grad = src.grad = None
                if ref.requires_grad:
                    ref.grad = None
                for p in mod.parameters():
                    p.grad = None

            with torch.no_grad():
                f = timeit(fwd, a.iters)
            fb = timeit(fwd_bwd, a.iters)
            print(json.dumps(dict(module="MSDeformAttn", call=call, tuning=a.set or "default", batch=a.batch, queries=S, autocast_bf16=amp, fused_preprocessing=fused,
                                  fwd_ms=round(f, 3), fwd_bwd_ms=round(fb, 3), peak_mem_GB=round(torch.cuda.max_memory_allocated() / 2**30, 2))),
                  flush=True)
            torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    main()
