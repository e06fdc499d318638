#!/usr/bin/env python
"""MonoDETR training-step benchmark (BASELINE.json configs[3]).

Drives the UNMODIFIED reference model + criterion + optimizer (staged by tools/stage_reference.py
into git-ignored baseline/_ref/MonoDETR) with this repo's MSDA op swapped in, on synthetic
KITTI-shaped batches, data-parallel with DistributedDataParallel (NCCL) when launched by torchrun.

    python tools/train_step_bench.py --steps 50 --warmup 10                      # 1 GPU
    torchrun --nproc-per-node 8 tools/train_step_bench.py --steps 50 --warmup 10  # 8 GPUs, 16 img/GPU
    python tools/train_step_bench.py --op ref_cuda      # same model, the reference's own CUDA kernels
                                                        # (oracle/_ref) behind the reference's own ops/ Python
    python tools/train_step_bench.py --host-opt all     # matcher / DDN targets / AdamW on the device
                                                        # (monosowa_b200.step_host); prints host_opt_check = loss terms
                                                        # of the patched vs the reference criterion on the same outputs
    python tools/train_step_bench.py --breakdown --profile-msda   # per-phase host/GPU times, criterion parts, top kernels
    ... --ddp default                                   # DDP as constructed by default (lean = static_graph, no buffer
                                                        # broadcast, is the harness default)

The step mirrors lib/helpers/trainer_helper.py:116-178: zero_grad, forward, SetCriterion, weighted sum,
reduce_dict + per-key .item() logging (``--logging faithful``; ``lean`` logs every 30th step only),
backward, the reference's own AdamW.  No reference file is edited; torch>=2 incompatibilities are
handled with import shims.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "MonoDETR")
sys.path.insert(0, ROOT)


def install_shims(op: str):
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} missing: run tools/stage_reference.py where /root/reference exists")
    sys.path.insert(0, REF)
    # modules the model file imports but never uses on this path
    for name in ("open3d",):
        sys.modules.setdefault(name, types.ModuleType(name))
    import torch.nn.modules.linear as _lin
    if not hasattr(_lin, "_LinearWithBias"):                       # ops/modules/ms_deform_attn.py:34
        _lin._LinearWithBias = _lin.NonDynamicallyQuantizableLinear
    if "torch._overrides" not in sys.modules:                       # ops/modules/ms_deform_attn.py:55
        import torch.overrides as _ov
        fake = types.ModuleType("torch._overrides")
        fake.has_torch_function, fake.handle_torch_function = _ov.has_torch_function, _ov.handle_torch_function
        sys.modules["torch._overrides"] = fake
    if op == "ours":
        import monosowa_b200.ops as b200_ops
        sys.modules["lib.models.monodetr.ops"] = b200_ops
        sys.modules["lib.models.monodetr.ops.modules"] = b200_ops.modules
        sys.modules["lib.models.monodetr.ops.functions"] = b200_ops.functions
    elif op == "ref_cuda":
        # the reference's own ops/ Python on top of the reference's own kernels (oracle/_ref)
        from oracle import msda_oracle as O
        ext = types.ModuleType("MultiScaleDeformableAttention")
        ext.ms_deform_attn_forward = lambda v, sh, lsi, loc, aw, step: O.ref_cuda_forward(v, sh, lsi, loc, aw)
        ext.ms_deform_attn_backward = lambda v, sh, lsi, loc, aw, g, step: O.ref_cuda_backward(v, sh, lsi, loc, aw, g.contiguous())
        sys.modules["MultiScaleDeformableAttention"] = ext
    else:
        raise ValueError(op)


def synthetic_batch(batch, device, seed, n_obj=8, max_objs=50):
    """N(0,1) images (B,3,384,1280), P2 calib, 8 synthetic Car labels/image."""
    g = torch.Generator().manual_seed(seed)
    U = lambda lo, hi, *shape: torch.rand(*shape, generator=g) * (hi - lo) + lo
    images = torch.randn(batch, 3, 384, 1280, generator=g)
    P2 = torch.tensor([[721.5, 0.0, 609.6, 44.9], [0.0, 721.5, 172.9, 0.2], [0.0, 0.0, 1.0, 0.003]])
    calibs = P2[None].repeat(batch, 1, 1)
    t = {
        "calibs": P2[None, None].repeat(batch, max_objs, 1, 1),
        "img_size": torch.tensor([[1242.0, 375.0]]).repeat(batch, 1),
        "labels": torch.zeros(batch, max_objs, dtype=torch.int8),
        "boxes": torch.zeros(batch, max_objs, 4), "boxes_3d": torch.zeros(batch, max_objs, 6),
        "depth": torch.zeros(batch, max_objs, 1), "size_3d": torch.zeros(batch, max_objs, 3),
        "heading_bin": torch.zeros(batch, max_objs, 1, dtype=torch.int64),
        "heading_res": torch.zeros(batch, max_objs, 1), "mask_2d": torch.zeros(batch, max_objs, dtype=torch.bool),
    }
    cx, cy = U(0.1, 0.9, batch, n_obj), U(0.4, 0.8, batch, n_obj)
    l, r = U(0.02, 0.08, batch, n_obj), U(0.02, 0.08, batch, n_obj)
    tp, b = U(0.02, 0.1, batch, n_obj), U(0.02, 0.1, batch, n_obj)
    t["boxes_3d"][:, :n_obj] = torch.stack([cx, cy, l, r, tp, b], -1)
    t["boxes"][:, :n_obj] = torch.stack([cx + (r - l) / 2, cy + (b - tp) / 2, l + r, tp + b], -1)
    t["depth"][:, :n_obj, 0] = U(5.0, 60.0, batch, n_obj)
    t["size_3d"][:, :n_obj] = torch.tensor([1.5, 1.6, 3.9]) * U(0.9, 1.1, batch, n_obj, 3)
    t["heading_bin"][:, :n_obj, 0] = torch.randint(0, 12, (batch, n_obj), generator=g)
    t["heading_res"][:, :n_obj, 0] = U(-0.26, 0.26, batch, n_obj)
    t["labels"][:, :n_obj] = 1
    t["mask_2d"][:, :n_obj] = True
    return images.to(device), calibs.to(device), {k: v.to(device) for k, v in t.items()}


def prepare_targets(targets, batch_size):                 # trainer_helper.py:180-191
    keys = ["labels", "boxes", "calibs", "depth", "size_3d", "heading_bin", "heading_res", "boxes_3d"]
    mask = targets["mask_2d"]
    return [{k: targets[k][i][mask[i]] for k in keys} for i in range(batch_size)]


def make_parser():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--op", default="ours", choices=["ours", "ref_cuda"])
    ap.add_argument("--logging", default="faithful", choices=["faithful", "lean"])
    ap.add_argument("--profile-msda", action="store_true", help="kineto pass: MSDA kernels' share of the step")
    ap.add_argument("--host-opt", default="none",
                    help="SURVEY 8 f3/f4: comma list of matcher,ddn,adamw,sdpa or 'all' -- device-resident replacements of the "
                         "step's host sections and fused SDPA for the decoder's dense attentions (monosowa_b200.step_host); "
                         "'none' = the reference's own code")
    ap.add_argument("--ddp", default="lean", choices=["default", "lean"],
                    help="lean: static_graph (the unused-parameter search runs once, not every step) and no per-step "
                         "buffer broadcast (the only buffers are FrozenBatchNorm statistics, which never change)")
    ap.add_argument("--breakdown", action="store_true",
                    help="extra pass: per-phase host issue time vs GPU time (forward / criterion / logging / backward / optimizer)")
    ap.add_argument("--mode", default="train", choices=["train", "infer"],
                    help="infer: model.eval() forward only (50 queries), as tester_helper.py:80-99")
    ap.add_argument("--amp", default="none", choices=["none", "bf16"],
                    help="bf16: autocast around the model forward (beyond the reference, which trains in fp32)")
    ap.add_argument("--no-fuse", action="store_true", help="parity mode: MSDeformAttn.fuse_preprocessing = False (literal path)")
    ap.add_argument("--dump-step", default="",
                    help="parity mode (tests/test_train_step_gpu.py): run ONE deterministic training step (seeded, dropout "
                         "off, no optimizer step) and save the loss terms and a sample of parameter gradients to this file")
    return ap


def default_args(**over):
    """argparse defaults as a namespace (bench.py calls run() without a command line)"""
    args = make_parser().parse_args([])
    for k, v in over.items():
        setattr(args, k, v)
    return args


def run(args):
    """One measurement; returns the result dict on rank 0 (None elsewhere).  Uses the default process group if one is
    initialised (bench.py under torchrun) and creates it otherwise when WORLD_SIZE > 1."""
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    own_pg = False
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=dev)
            own_pg = True
    os.environ.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or 8) // max(world, 1))))
    torch.set_num_threads(max(1, (os.cpu_count() or 8) // max(world, 1)))

    install_shims(args.op)
    import yaml
    from lib.models.monodetr import build as build_monodetr
    from lib.helpers.optimizer_helper import build_optimizer
    from utils import misc
    cfg = yaml.load(open(os.path.join(REF, "configs", "monodetr.yaml")), Loader=yaml.Loader)
    cfg["model"]["pretrained"] = False                      # no network for torchvision weights
    torch.manual_seed(cfg.get("random_seed", 444))
    model, criterion = build_monodetr(cfg["model"])
    model.to(dev).train()
    criterion.train()
    n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
    optimizer = build_optimizer(cfg["optimizer"], model)
    images, calibs, tdict = synthetic_batch(args.batch, dev, seed=1237 + rank)
    targets = prepare_targets(tdict, args.batch)
    host_opt, host_opt_check = [], None
    if args.host_opt != "none":
        from monosowa_b200 import step_host
        # "all" = what pays: fused SDPA for the dense attentions is an opt-in ("all,sdpa"): measured SLOWER in fp32 on B200
        # (PyTorch's only fp32 fused kernel is the sm80 SIMT memory-efficient one: 24 vs ~5 ms per step, profiles/r02_training_step.md)
        want = set(args.host_opt.split(","))
        if "all" in want:
            want |= {"matcher", "ddn", "adamw", "bn"}
        with torch.no_grad():                               # same model outputs through the reference criterion ...
            torch.manual_seed(4321)                         # (same dropout masks for the forward after the patches)
            out0 = model(images, calibs, targets, tdict["img_size"], dn_args=None)
            ld_ref = {k: float(v) for k, v in criterion(out0, targets, None, None).items()}
        host_opt = step_host.install(criterion, optimizer, matcher="matcher" in want, ddn="ddn" in want, adamw="adamw" in want,
                                     model=model, attention="sdpa" in want, frozen_bn="bn" in want)
        with torch.no_grad():                               # ... and through the patched one: every loss term must agree
            ld_opt = {k: float(v) for k, v in criterion(out0, targets, None, None).items()}
        host_opt_check = {"loss_terms": len(ld_ref), "max_abs_diff": max(abs(ld_ref[k] - ld_opt[k]) for k in ld_ref),
                          "max_rel_diff": max(abs(ld_ref[k] - ld_opt[k]) / max(abs(ld_ref[k]), 1e-12) for k in ld_ref)}
        if "frozen_bn" in host_opt or "sdpa" in host_opt:   # patches inside the model: its outputs before / after
            with torch.no_grad():
                torch.manual_seed(4321)
                out1 = model(images, calibs, targets, tdict["img_size"], dn_args=None)
            same = [bool(torch.equal(out0[k], out1[k])) for k in out0 if isinstance(out0[k], torch.Tensor)]
            host_opt_check["model_outputs_bitwise_equal"] = all(same)
            host_opt_check["model_outputs_compared"] = len(same)
            del out1
        del out0
    img_sizes = tdict["img_size"]
    weight_dict = criterion.weight_dict

    if args.dump_step:
        # One deterministic step for the ours-vs-reference-kernels parity test: dropout off (the two op
        # implementations must see the same network), fixed seed, gradients of a fixed sample of parameters.
        for mod in model.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
            if args.no_fuse and hasattr(mod, "fuse_preprocessing"):
                mod.fuse_preprocessing = False
        torch.manual_seed(1234)
        optimizer.zero_grad()
        outputs = model(images, calibs, targets, img_sizes, dn_args=None)
        ld = criterion(outputs, targets, None, None)
        loss = sum(ld[k] * weight_dict[k] for k in ld.keys() if k in weight_dict)
        loss.backward()
        torch.cuda.synchronize()
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad and p.grad is not None]
        pick = [named[i] for i in range(0, len(named), max(1, len(named) // 48))]
        torch.save({"loss": float(loss), "loss_terms": {k: float(v) for k, v in ld.items()},
                    "grads": {n: p.grad.detach().float().cpu() for n, p in pick},
                    "grad_norm_all": float(torch.sqrt(sum(p.grad.double().pow(2).sum() for _, p in named))),
                    "op": args.op, "host_opt": host_opt, "n_params_with_grad": len(named)}, args.dump_step)
        return {"dumped": args.dump_step, "loss": float(loss), "op": args.op}

    net = model
    if world > 1:
        # one dry step without DDP to learn whether every parameter receives a gradient
        out = model(images, calibs, targets, img_sizes, dn_args=None)
        ld = criterion(out, targets, None, None)
        sum(ld[k] * weight_dict[k] for k in ld if k in weight_dict).backward()
        unused = any(p.requires_grad and p.grad is None for p in model.parameters())
        optimizer.zero_grad()
        lean = args.ddp == "lean"
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], find_unused_parameters=unused,
                                                        gradient_as_bucket_view=True, static_graph=lean,
                                                        broadcast_buffers=not lean)
        if rank == 0:
            print(f"[ddp] unused parameters: {unused}; mode {args.ddp}", file=sys.stderr)

    amp = torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.amp == "bf16")

    def to_f32(o):
        if isinstance(o, torch.Tensor):
            return o.float() if o.is_floating_point() else o
        if isinstance(o, dict):
            return {k: to_f32(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return type(o)(to_f32(v) for v in o)
        return o

    if args.mode == "infer":
        model.eval()

        def step(i):                                             # tester_helper.py:94-99
            with torch.no_grad(), amp:
                outputs = net(images, calibs, targets, img_sizes, dn_args=None)
            return outputs["pred_logits"].float().sum()
    else:
        step = None

    def train_step(i):
        optimizer.zero_grad()
        with amp:
            outputs = net(images, calibs, targets, img_sizes, dn_args=None)
        outputs = to_f32(outputs)
        ld = criterion(outputs, targets, None, None)
        loss = sum(ld[k] * weight_dict[k] for k in ld.keys() if k in weight_dict)
        if args.logging == "faithful" or i % 30 == 0:           # trainer_helper.py:150-158
            red = misc.reduce_dict(ld)
            _ = sum((red[k] * weight_dict[k]).item() for k in red if k in weight_dict)
        loss.backward()
        optimizer.step()
        return loss

    if step is None:
        step = train_step

    for i in range(args.warmup):
        loss = step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t0 = time.perf_counter()
    ev[0].record()
    for i in range(args.steps):
        loss = step(i + 1)
        ev[i + 1].record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    ms = torch.tensor([statistics.median(per), sum(per) / len(per), wall * 1e3 / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)

    breakdown = None
    if args.breakdown and args.mode == "train" and rank == 0 and world == 1:
        # each phase bracketed by CUDA events (GPU time between the phase's first and last kernel incl. idle gaps) and by
        # perf_counter without a sync (how long the host needs to ISSUE the phase); a phase whose issue time exceeds
        # its GPU time is host-bound.  `sync_ms` = the same phase followed by a device sync (max of the two).
        names = ("zero_grad", "forward", "criterion", "logging", "backward", "optimizer")
        acc = {n: [0.0, 0.0] for n in names}
        n_bd = 5
        for i in range(n_bd):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
            ts = []
            torch.cuda.synchronize()
            evs[0].record(); ts.append(time.perf_counter())
            optimizer.zero_grad()
            evs[1].record(); ts.append(time.perf_counter())
            with amp:
                outputs = net(images, calibs, targets, img_sizes, dn_args=None)
            outputs = to_f32(outputs)
            evs[2].record(); ts.append(time.perf_counter())
            ld = criterion(outputs, targets, None, None)
            loss = sum(ld[k] * weight_dict[k] for k in ld.keys() if k in weight_dict)
            evs[3].record(); ts.append(time.perf_counter())
            red = misc.reduce_dict(ld)
            _ = sum((red[k] * weight_dict[k]).item() for k in red if k in weight_dict)
            evs[4].record(); ts.append(time.perf_counter())
            loss.backward()
            evs[5].record(); ts.append(time.perf_counter())
            optimizer.step()
            evs[6].record(); ts.append(time.perf_counter())
            torch.cuda.synchronize()
            for j, n in enumerate(names):
                acc[n][0] += (ts[j + 1] - ts[j]) * 1e3 / n_bd
                acc[n][1] += evs[j].elapsed_time(evs[j + 1]) / n_bd
        breakdown = {n: {"host_issue_ms": round(v[0], 2), "gpu_span_ms": round(v[1], 2)} for n, v in acc.items()}
        # inside the criterion: wall time (device-synchronised) of the matcher calls and of each loss term
        crit_t = {}

        def timed(label, fn):
            def wrapper(*a, **k):
                torch.cuda.synchronize(); t0_ = time.perf_counter()
                r = fn(*a, **k)
                torch.cuda.synchronize()
                crit_t[label] = crit_t.get(label, 0.0) + (time.perf_counter() - t0_) * 1e3 / n_bd
                return r
            return wrapper

        orig_matcher, orig_get_loss = criterion.matcher.forward, criterion.get_loss
        criterion.matcher.forward = timed("matcher(x3)", orig_matcher)
        criterion.get_loss = lambda loss, *a, **k: timed("loss:" + loss, orig_get_loss)(loss, *a, **k)
        with torch.no_grad():
            for i in range(n_bd):
                with amp:
                    outputs = net(images, calibs, targets, img_sizes, dn_args=None)
                criterion(to_f32(outputs), targets, None, None)
        criterion.matcher.forward, criterion.get_loss = orig_matcher, orig_get_loss
        breakdown["criterion_parts_ms"] = {k: round(v, 2) for k, v in sorted(crit_t.items(), key=lambda kv: -kv[1])}

    share = None
    if args.profile_msda and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for i in range(3):
                step(i + 1)
            torch.cuda.synchronize()
        tot = msda_t = 0.0
        for e in prof.key_averages():
            dt = getattr(e, "device_time_total", 0.0) or getattr(e, "cuda_time_total", 0.0)
            if e.device_type == torch.autograd.DeviceType.CUDA:
                tot += dt
                if "msda::" in e.key or "ms_deformable" in e.key or "rec_kernel" in e.key or "vec_kernel" in e.key:
                    msda_t += dt
        top = sorted(((getattr(e, "device_time_total", 0.0) or getattr(e, "cuda_time_total", 0.0), e.count, e.key)
                      for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA), reverse=True)[:25]
        share = {"gpu_kernel_ms_per_step": tot / 3e3, "msda_kernel_ms_per_step": msda_t / 3e3,
                 "msda_share_of_gpu_time": msda_t / tot if tot else None,
                 "top_kernels_ms_per_step": [[round(t / 3e3, 3), c // 3, k[:90]] for t, c, k in top]}
    elif args.profile_msda and world > 1:
        for i in range(3):
            step(i + 1)                                          # keep the ranks in lock step

    result = None
    if rank == 0:
        med, mean, wallms = ms.tolist()
        result = {
            "metric": "MonoDETR train img/s" if args.mode == "train" else "MonoDETR inference img/s", "value": args.batch * world / (mean * 1e-3), "unit": "img/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean, "ms_per_step_median": med,
            "ms_per_step_wall": wallms, "scaling": "weak", "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[3]: unmodified reference MonoDETR (ResNet-50, 3 enc + 3 dec layers) + "
                                   "SetCriterion + reference AdamW, synthetic KITTI batch", "batch_per_gpu": args.batch,
                       "global_batch": args.batch * world, "image": [384, 1280], "msda_op": args.op, "host_opt": host_opt, "host_opt_check": host_opt_check, "logging": args.logging, "mode": args.mode, "amp": args.amp,
                       "parallelism": f"ddp{world}", "ddp": args.ddp, "trainable_params": n_params,
                       "allreduce_bytes_per_step": 4 * n_params if world > 1 else 0},
            "loss": float(loss), "msda": share, "breakdown": breakdown}
    del net, model, criterion, optimizer
    torch.cuda.empty_cache()
    if own_pg:
        dist.destroy_process_group()
    return result


def main():
    args = make_parser().parse_args()
    result = run(args)
    if result is not None:
        print(json.dumps(result), flush=True)


if __name__ == "__main__":
    main()
