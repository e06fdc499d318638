#!/usr/bin/env python
"""bench.py -- MSDA forward+backward throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (oracle port), rank 0 only
    torchrun ... bench.py --gpus N ...                       # one rank per GPU (weak scaling by batch)

One "step" = one MSDA forward + backward over one synthetic batch of the headline workload
(BASELINE.json configs[1]: KITTI 384x1280 encoder self-attention, 4 levels 48x160..6x20,
10200 queries, 8 heads x 32 channels, 4 points, batch 16, fp32).  The op is batch-local
(SURVEY.md 8e): with N GPUs every rank runs its own batch of 16, no data-path collective.

Printed JSON (rank 0, one line):
  value      algorithmic GB moved per second, whole job, inputs resident in HBM
             (algorithmic bytes: SURVEY.md 8d / monosowa_b200.workloads.algorithmic_bytes)
  e2e        same metric through the public API (MSDeformAttnFunction.apply + autograd) with
             pinned HOST buffers: H2D of value/loc/attn/grad_out and D2H of out + the three
             gradients inside the timed region
  roofline   dominant kernel (the backward scatter): algorithmic bytes / CUDA-event time vs the
             measured HBM copy peak in MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the reference's ms_deform_attn_core_pytorch on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "MSDA fwd+bwd GB/s"
UNIT = "GB/s"
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback


# ------------------------------------------------------------------------------------------
# helpers (importable by the CPU tests)
# ------------------------------------------------------------------------------------------
def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def aggregate(local_bytes_per_step: float, local_ms_per_step: float, world: int, reduce_fn=None):
    """Whole-job throughput: all ranks' bytes / the slowest rank's time.  `reduce_fn(tensor, op)`
    performs the all-reduce (None = single process)."""
    t = torch.tensor([local_ms_per_step], dtype=torch.float64)
    b = torch.tensor([local_bytes_per_step], dtype=torch.float64)
    if reduce_fn is not None and world > 1:
        t = reduce_fn(t, "max")
        b = reduce_fn(b, "sum")
    ms = float(t.item())
    return float(b.item()) / (ms * 1e-3) / 1e9, ms


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc, self.gpu = None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
            out = ""
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path (ms_deform_attn_core_pytorch)
# ------------------------------------------------------------------------------------------
def load_workloads():
    """monosowa_b200/workloads.py as a stand-alone module: pure tensor plumbing, and importing it this way does NOT
    import the package, i.e. does not load libmsda_b200.so -- the reference arm must not touch the product."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_msda_workloads", os.path.join(ROOT, "monosowa_b200", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_msda_workloads"] = mod
    spec.loader.exec_module(mod)
    return mod


REF_FUNC_FILE = os.path.join(ROOT, "baseline", "_ref", "MonoDETR", "lib", "models", "monodetr", "ops", "functions",
                             "ms_deform_attn_func.py")


def load_reference_core():
    """The UNMODIFIED reference function ms_deform_attn_core_pytorch (ops/functions/ms_deform_attn_func.py:41-61) from
    the staged reference tree (baseline/_ref, tools/stage_reference.py).  The file imports the compiled extension at
    line 18; an empty stand-in module satisfies that import (the function itself never touches it).  None if the tree
    is not staged."""
    if not os.path.exists(REF_FUNC_FILE):
        return None
    import importlib.util
    import types
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    spec = importlib.util.spec_from_file_location("_ref_ms_deform_attn_func", REF_FUNC_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ms_deform_attn_core_pytorch


def cpu_reference_step_fn(wl, sample_batch, W=None):
    """Returns (step, bytes_per_step, description, kind).  The sample is `sample_batch` images of the
    workload -- the grid_sample path materialises (N*M, D, Lq, L*P) and needs ~0.1 s per image."""
    W = W or load_workloads()
    swl = W.config(1, batch=sample_batch, loc_mode=wl.loc_mode, seed=wl.seed)
    d = W.make_inputs(swl, device="cpu")
    core = load_reference_core()
    kind = "reference"
    what = "the reference's own ms_deform_attn_core_pytorch (baseline/_ref, unmodified) forward + autograd backward"
    if core is None:                                         # tree not staged: the oracle's restatement of the same function
        from oracle import msda_oracle as O                  # checker used as the CPU baseline only
        core, kind = O.core_grid_sample, "port"
        what = "oracle.core_grid_sample (restated ms_deform_attn_core_pytorch, grid_sample) forward + autograd backward"

    def step():
        v = d["value"].detach().clone().requires_grad_(True)
        l = d["loc"].detach().clone().requires_grad_(True)
        a = d["attn"].detach().clone().requires_grad_(True)
        out = core(v, d["shapes"], l, a)
        out.backward(d["grad_out"].reshape(out.shape))
        return out

    return step, W.algorithmic_bytes(swl)["total"], f"{sample_batch} of {wl.batch} images of {wl.name} per step", kind, what


def time_cpu(step, steps, warmup):
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return statistics.mean(ts) * 1e3


def run_reference_arm(args, wl):
    rank, _, world = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, nbytes, sample, kind, what = cpu_reference_step_fn(wl, sample_batch=2)
    steps = max(1, min(args.steps, 8))
    ms = time_cpu(step, steps, max(1, min(args.warmup, 2)))
    val = nbytes / (ms * 1e-3) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": max(1, min(args.warmup, 2)), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, world),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample,
                         "what": what},
        "native_libraries_loaded": [m for m in ("monosowa_b200", "monosowa_b200._lib") if m in sys.modules],
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(wl, world):
    return {"workload": f"BASELINE.json configs[1]: MSDA encoder self-attention fwd+bwd, KITTI 384x1280 "
                        f"(levels {wl.shapes}), {wl.Lq} queries, {wl.heads} heads x {wl.head_dim}, {wl.points} points, "
                        f"batch {wl.batch}/GPU, {str(wl.dtype).replace('torch.', '')}",
            "name": wl.name, "batch_per_gpu": wl.batch, "global_batch": wl.batch * world, "queries": wl.Lq,
            "loc_mode": wl.loc_mode, "parallelism": f"replicas x{world} (batch-sharded, no collective)",
            "l2_policy": "inputs (585 MB/step) exceed the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------
def run_ours(args, wl):
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import monosowa_b200 as msda
    from monosowa_b200 import workloads as W
    for kv in filter(None, args.tune.split(",")):
        k_, v_ = kv.split("=")
        msda._lib.set_tuning(k_, int(v_))
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    d = W.make_inputs(wl, device=dev, seed=wl.seed + 1000 * rank)
    ab = W.algorithmic_bytes(wl)
    fwd_op, bwd_op = torch.ops.msda.forward, torch.ops.msda.backward
    a5 = (d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"])

    def step_device():
        out = fwd_op(*a5, 64)
        return out, bwd_op(*a5, d["grad_out"], 64)

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) + per-kernel events (roofline) ----------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                     # started before warm-up: nvidia-smi needs ~0.1 s to come up
    # Warm-up keeps the previous step's outputs alive while the next step allocates, exactly like the
    # timed loop below, so the caching allocator has reached its steady state (no cudaMalloc inside
    # the timed region); the 0.3 s pre-warm lets clocks settle before the W warm-up steps.
    t_pre = time.perf_counter()
    keep = None
    while time.perf_counter() - t_pre < 0.3:
        keep = step_device()
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        keep = step_device()
    del keep
    barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    n0 = msda._lib.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    out = grads = None
    for i in range(args.steps):
        ev[i][0].record()
        out = fwd_op(*a5, 64)
        ev[i][1].record()
        grads = bwd_op(*a5, d["grad_out"], 64)
        ev[i][2].record()
    t_end.record()
    barrier()
    launches = msda._lib.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = t_start.elapsed_time(t_end) / args.steps
    fwd_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in ev)
    bwd_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in ev)
    if os.environ.get("MSDA_BENCH_TRACE") and rank == 0:       # per-step trace for drift / throttling diagnosis
        with open(os.environ["MSDA_BENCH_TRACE"], "w") as f:
            json.dump({"fwd_ms": [e[0].elapsed_time(e[1]) for e in ev], "bwd_ms": [e[1].elapsed_time(e[2]) for e in ev]}, f)

    def reduce_fn(t, op):
        tt = t.to(dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return tt.cpu()

    value, ms_max = aggregate(ab["total"], ms_step, world, reduce_fn if use_dist else None)

    # sustained figure: the timed region above lasts tens of milliseconds at the driver's K; this block runs
    # `--sustain-steps` more steps back to back (one event pair, >= 0.2 s) so that a power-capped clock shows
    barrier()
    for _ in range(3):                      # the GPU sat idle while rank 0 stopped the clock sampler: not part of "sustained"
        keep = step_device()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(args.sustain_steps):
        keep = step_device()
    s1.record()
    barrier()
    del keep
    value_sustained, ms_sustained = aggregate(ab["total"], s0.elapsed_time(s1) / max(args.sustain_steps, 1), world,
                                              reduce_fn if use_dist else None)
    live_rows = W.live_corner_rows(d, wl)                    # gathered (and scattered) 128-byte rows that carry weight

    # ---- end to end: pinned host buffers in, results back to pinned host buffers -------------
    # Device staging buffers and pinned result buffers are allocated once; every step copies all
    # inputs host->device and all four results device->host, chunked over the batch on three
    # streams (H2D / compute / D2H) so that PCIe in, the kernels and PCIe out overlap.
    host_in = {k: d[k].cpu().pin_memory() for k in ("value", "loc", "attn", "grad_out")}
    h2d_bytes = sum(t.numel() * t.element_size() for t in host_in.values())
    dev_in = {k: torch.empty_like(d[k]) for k in host_in}
    host_out = [torch.empty(s_, dtype=dt_).pin_memory() for s_, dt_ in
                ((tuple(out.shape), out.dtype), (tuple(d["value"].shape), d["value"].dtype),
                 (tuple(d["loc"].shape), d["loc"].dtype), (tuple(d["attn"].shape), d["attn"].dtype))]
    F = msda.MSDeformAttnFunction.apply
    chunks = args.e2e_chunks if wl.batch % args.e2e_chunks == 0 else 1
    cb = wl.batch // chunks
    s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()

    def step_e2e():
        cur = torch.cuda.current_stream()
        for s in (s_in, s_cmp, s_out):
            s.wait_stream(cur)
        for c in range(chunks):
            sl = slice(c * cb, (c + 1) * cb)
            with torch.cuda.stream(s_in):
                for k in host_in:
                    dev_in[k][sl].copy_(host_in[k][sl], non_blocking=True)
                e_in = torch.cuda.Event(); e_in.record()
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(e_in)
                v = dev_in["value"][sl].detach().requires_grad_(True)
                l = dev_in["loc"][sl].detach().requires_grad_(True)
                a = dev_in["attn"][sl].detach().requires_grad_(True)
                o = F(v, d["shapes"], d["lsi"], l, a, 64)
                o.backward(dev_in["grad_out"][sl])
                e_c = torch.cuda.Event(); e_c.record()
            with torch.cuda.stream(s_out):
                s_out.wait_event(e_c)
                for h, t in zip(host_out, (o.detach(), v.grad, l.grad, a.grad)):
                    h[sl].copy_(t, non_blocking=True)
                    t.record_stream(s_out)
        cur.wait_stream(s_out)
        cur.wait_stream(s_in)

    def time_e2e(step_fn):
        for _ in range(max(1, min(args.warmup, 3))):
            step_fn()
        barrier()
        n_ = max(1, min(args.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_):
            step_fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / n_

    # (a) the host-buffer entry point of the C ABI (msda_host_step_f32): one call per step, the library
    #     pipelines the batch image by image on its own copy streams
    go_host = host_in["grad_out"].view(wl.batch, -1, wl.heads * wl.head_dim)

    def step_host():
        msda.host_step(host_in["value"], d["shapes"], d["lsi"], host_in["loc"], host_in["attn"], go_host,
                       images_per_chunk=args.e2e_images_per_chunk, results=tuple(host_out), synchronize=False,
                       stages=args.e2e_stages)

    e2e_ms = time_e2e(step_host)
    chk = torch.equal(host_out[0].to(dev), out) if rank == 0 else True
    # (b) the same through MSDeformAttnFunction.apply + autograd with a Python-level 3-stream pipeline
    e2e_autograd_ms = time_e2e(step_e2e)
    d2h_bytes = sum(t.numel() * t.element_size() for t in host_out)

    # PCIe copy rates of this box (the e2e leg's own roofline): the same pinned buffers, one direction at a
    # time and both at once; lower bound of an e2e step = max(h2d_bytes / h2d rate, d2h_bytes / d2h rate), duplex
    def copy_rate(h2d, d2h, reps=3):
        res = (out, grads[0], grads[1], grads[2])
        best = float("inf")
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if h2d:
                with torch.cuda.stream(s_in):
                    for k in host_in:
                        dev_in[k].copy_(host_in[k], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s_out):
                    for h, t in zip(host_out, res):
                        h.copy_(t, non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best
    # all ranks copy at the same time (barrier in front of every probe), slowest rank counts: at N > 1 this is what the
    # box's host memory / PCIe fabric gives N GPUs together, i.e. the floor the N-GPU e2e figure has to be read against
    def probe(h2d, d2h):
        barrier()
        t = torch.tensor([copy_rate(h2d, d2h)], dtype=torch.float64)
        return float((reduce_fn(t, "max") if use_dist else t).item())
    t_h, t_d, t_b = probe(True, False), probe(False, True), probe(True, True)
    pcie = {"h2d_GBps": h2d_bytes * world / t_h / 1e9, "d2h_GBps": d2h_bytes * world / t_d / 1e9,
            "duplex_ms": t_b * 1e3, "ranks_copying_at_once": world,
            "note": "full-duplex copy of one step's inputs and results on every rank at once, no kernels: the floor of an "
                    "e2e step on this box (whole-job GB/s per direction; duplex_ms = slowest rank)"}
    e2e_value, e2e_ms_max = aggregate(ab["total"], e2e_ms, world, reduce_fn if use_dist else None)
    chk = chk and (torch.equal(host_out[0].to(dev), out) if rank == 0 else True)   # both e2e paths reproduce the device forward

    ref_cuda = time_reference_cuda(d, ab, args) if (world == 1 and rank == 0 and not args.no_ref_cuda) else None
    # configs[2] / configs[4] on every rank (configs[4] is "at 8 B200": one replica per GPU, no collective in the data
    # path); per-row times are the slowest rank's, GB/s is the whole job's
    other = None
    if not args.no_other_configs:
        try:
            other = other_configs_table(msda, W, dev)
            if use_dist:
                t = torch.tensor([[r["fwd_us"], r["bwd_us"]] for r in other], dtype=torch.float64)
                t = reduce_fn(t, "max")
                for r, (f_us, b_us) in zip(other, t.tolist()):
                    r["GBps"] = round(r["GBps"] * (r["fwd_us"] + r["bwd_us"]) / (f_us + b_us) * world, 1)
                    r["fwd_us"], r["bwd_us"], r["replicas"] = round(f_us, 1), round(b_us, 1), world
        except Exception as exc:  # noqa: BLE001
            other = [{"unavailable": repr(exc)[:200]}]

    # ---- the second half of BASELINE.json's metric: the MonoDETR training step (configs[3]) on the same ranks ------
    kernels = {"fwd": msda._lib.describe("forward", wl.dtype, wl.batch, wl.heads, wl.head_dim, wl.L, wl.points, wl.Lq),
               "bwd": msda._lib.describe("backward", wl.dtype, wl.batch, wl.heads, wl.head_dim, wl.L, wl.points, wl.Lq)}
    out_shape_checked = bool(chk)
    del d, a5, out, grads, host_in, dev_in, host_out, go_host, step_device, step_e2e, step_host, fwd_op, bwd_op, copy_rate, time_e2e
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    train = None if args.no_train_step else train_step_section(args, rank, world)

    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    dom = ("bwd", bwd_ms, ab["bwd"]) if bwd_ms >= fwd_ms else ("fwd", fwd_ms, ab["fwd"])
    achieved = dom[2] / (dom[1] * 1e-3) / 1e9
    ceiling = binding_ceiling(live_rows, load_traffic("bwd_red_lines_per_image"), wl.batch)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {torch.float32: "f32", torch.bfloat16: "bf16", torch.float64: "f64"}[wl.dtype], "data": "synthetic",
        "config": workload_config(wl, world),
        "value_sustained": value_sustained, "ms_per_step_sustained": ms_sustained, "sustain_steps": args.sustain_steps,
        "frac_of_hbm_peak": value / world / peak,
        "frac_of_nominal_hbm_peak": value / world / 8000.0,      # north_star's nominal 8 TB/s, for reference
        "fwd_ms": fwd_ms, "bwd_ms": bwd_ms,
        "fwd_GBps": ab["fwd"] / (fwd_ms * 1e-3) / 1e9, "bwd_GBps": ab["bwd"] / (bwd_ms * 1e-3) / 1e9,
        "algorithmic_bytes": {"fwd": ab["fwd"], "bwd": ab["bwd"], "gather_cache_level": ab["gather"]},
        "kernels": kernels,
        "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": load_traffic(dom[0]), "peak_source": peak_src,
                     "note": "achieved = algorithmic bytes of the launch / CUDA-event time on the launch stream; "
                             "bwd includes its cudaMemsetAsync of grad_value",
                     # the HBM roofline is not the binding one for this gather/scatter:
                     "binding_unit": ceiling["binding_unit"], "binding_ceiling": ceiling,
                     "frac_of_binding_ceiling": ceiling["fwd_plus_bwd_floor_ms"] / (fwd_ms + bwd_ms),
                     "frac_of_binding_ceiling_dominant_kernel": ceiling[dom[0] + "_floor_ms"] / dom[1]},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms_max, "h2d_bytes_per_step": h2d_bytes * world,
                "d2h_bytes_per_step": d2h_bytes * world, "bytes_note": "whole job (all ranks), like `value`", "chunks": chunks, "matches_device_path": out_shape_checked, "pcie": pcie,
                "frac_of_pcie_floor": (pcie["duplex_ms"] / e2e_ms_max) if pcie else None,
                "api": "msda_host_step_f32 (C ABI, pinned host buffers in and out; monosowa_b200.host_step), "
                       f"{args.e2e_images_per_chunk} image(s) per pipeline chunk, {args.e2e_stages} stages",
                "autograd_api_ms_per_step": e2e_autograd_ms,
                "autograd_api": f"MSDeformAttnFunction.apply + autograd, Python 3-stream pipeline, {chunks} chunks"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "lib": msda._lib.build_info(),
        "train_step": train,
        "other_configs": other,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        step, nbytes, sample, kind, what = cpu_reference_step_fn(wl, sample_batch=2, W=W)
        ms = time_cpu(step, args.cpu_steps, 1)
        line["cpu_baseline"] = {"value": nbytes / (ms * 1e-3) / 1e9, "unit": UNIT, "cores": torch.get_num_threads(),
                                "kind": kind, "what": what, "sample": sample + f", {args.cpu_steps} timed steps", "ms_per_step": ms}
    if ref_cuda is not None:
        line["ref_cuda_kernel"] = ref_cuda
    print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()


def train_step_section(args, rank, world):
    """BASELINE.json configs[3] through tools/train_step_bench.py on the ranks this process group already has:
    the unmodified reference MonoDETR (staged under baseline/_ref) + SetCriterion + AdamW, DDP over NCCL, with this
    repo's op and device-resident host sections.  At one GPU the same step is also timed with the reference's own
    CUDA kernels (oracle/_ref) in a child process.  Returns a dict on rank 0 (None elsewhere); never raises."""
    try:
        from tools import train_step_bench as T
        if not os.path.isdir(T.REF):
            return {"unavailable": f"{T.REF} not staged (python tools/stage_reference.py where /root/reference exists)"} if rank == 0 else None
        res = T.run(T.default_args(steps=args.train_steps, warmup=args.train_warmup, batch=16, op="ours", host_opt="all",
                                   profile_msda=(world == 1), breakdown=(world == 1)))
        if rank != 0:
            return None
        out = {"metric": res["metric"], "value": res["value"], "unit": res["unit"], "n_gpus": res["n_gpus"],
               "ms_per_step": res["ms_per_step"], "ms_per_step_median": res["ms_per_step_median"], "steps": res["steps"],
               "warmup": res["warmup"], "batch_per_gpu": 16, "global_batch": 16 * world, "scaling": "weak",
               "parallelism": res["config"]["parallelism"], "ddp": res["config"]["ddp"],
               "allreduce_bytes_per_step": res["config"]["allreduce_bytes_per_step"],
               "trainable_params": res["config"]["trainable_params"], "host_opt": res["config"]["host_opt"],
               "host_opt_check": res["config"]["host_opt_check"], "loss": res["loss"], "dtype": "f32",
               "workload": res["config"]["workload"],
               "timing": "CUDA events around every step on the training stream, mean over the timed steps, max over ranks"}
        if res.get("msda"):
            out["msda_share_of_gpu_time"] = res["msda"]["msda_share_of_gpu_time"]
            out["msda_kernel_ms_per_step"] = res["msda"]["msda_kernel_ms_per_step"]
            out["gpu_kernel_ms_per_step"] = res["msda"]["gpu_kernel_ms_per_step"]
        if res.get("breakdown"):
            bd = res["breakdown"]
            out["breakdown"] = {k: v for k, v in bd.items() if k != "criterion_parts_ms"}
            # What limits the step: the share of the step during which the GPU executes kernels (kineto, 3 profiled steps).
            # Phases whose host issue time equals their GPU span end in a device synchronisation of the reference's own code
            # (num_boxes.item(), boolean indexing), so the equality alone does not tell who waits for whom.
            busy = res["msda"]["gpu_kernel_ms_per_step"] / res["ms_per_step"] if res.get("msda") else None
            out["limited_by"] = {
                "gpu_busy_fraction": busy,
                "verdict": ("GPU time" if busy is not None and busy >= 0.8 else "host issue time"),
                "note": "the reference model in its own precision: fp32 SIMT GEMMs of the Linears (TF32 off, the torch default) "
                        "and TF32 cuDNN convolutions of ResNet-50 make up most of the kernel time; MSDA is msda_share_of_gpu_time "
                        "of it.  Every rank runs the same step, DDP's 149.8 MB all-reduce overlaps the backward: weak scaling "
                        "is flat in ms_per_step"}
        if world == 1 and not args.no_ref_cuda:
            cmd = [sys.executable, os.path.join(ROOT, "tools", "train_step_bench.py"), "--op", "ref_cuda", "--steps",
                   str(max(5, args.train_steps // 2)), "--warmup", str(args.train_warmup)]
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
                ref = json.loads(r.stdout.strip().splitlines()[-1])
                out["ref_cuda_op"] = {"value": ref["value"], "unit": "img/s", "ms_per_step": ref["ms_per_step"],
                                      "what": "same model, criterion and optimizer with the reference's own ops/ Python on the "
                                              "reference's CUDA kernels recompiled for sm_100a (oracle/_ref), host sections as the reference has them"}
            except Exception as exc:  # noqa: BLE001
                out["ref_cuda_op"] = {"unavailable": repr(exc)[:200]}
        return out
    except Exception as exc:  # noqa: BLE001
        import traceback
        return {"unavailable": repr(exc)[:300], "trace": traceback.format_exc()[-600:]} if rank == 0 else None


# Measured unit rates behind the binding ceiling (B200, this pool; microbenchmarks under tools/ubench, outputs under
# profiles/): rows = 128-byte (pixel, head) rows of `value` / `grad_value`
L2_TO_SM_ROWS_PER_S = 20.6e12 / 128          # L2-resident random row gather, all SMs (r01_ubench_gather_rows.txt: 20.6 TB/s)
L1_PIPE_ROWS_PER_S = 148 * 1.965e9           # one row per clock per SM through the L1 data pipe (LDS.128 rows: 1.06 clk)
L2_RED_LINES_PER_S = 49.6e9                  # REDG.128 lines absorbed by L2 chip-wide (r01b_ubench_lines_per_instruction.txt)


def binding_ceiling(live_rows, red_lines_per_image, batch):
    """Attainable floor of forward and backward from the units that actually bind them:
    forward  = live corner rows / L2->SM row-fill rate (the rows do not fit L1; 0.245 ms more if they all did);
    backward = max(the same gather, the reduction lines its scatter sends to L2 / L2's reduction rate)."""
    red_lines = (red_lines_per_image or 0) * batch
    fwd = live_rows / L2_TO_SM_ROWS_PER_S * 1e3
    bwd_gather = live_rows / L2_TO_SM_ROWS_PER_S * 1e3
    bwd_red = red_lines / L2_RED_LINES_PER_S * 1e3
    return {"binding_unit": "forward: L2->SM row fills / L1 data pipe (one 128-byte row per sample corner); backward: "
                            "L2 reduction rate (REDG.128 lines) and the same gather",
            "live_rows_per_launch": live_rows, "bwd_reduction_lines_per_launch": red_lines,
            "rates": {"l2_to_sm_rows_per_s": L2_TO_SM_ROWS_PER_S, "l1_pipe_rows_per_s": L1_PIPE_ROWS_PER_S,
                      "l2_reduction_lines_per_s": L2_RED_LINES_PER_S},
            "fwd_floor_ms": fwd, "fwd_floor_if_all_rows_hit_l1_ms": live_rows / L1_PIPE_ROWS_PER_S * 1e3,
            "bwd_floor_ms": max(bwd_gather, bwd_red), "fwd_plus_bwd_floor_ms": fwd + max(bwd_gather, bwd_red),
            "source": "unit rates: tools/ubench microbenchmarks (profiles/r01_ubench_gather_rows.txt, "
                      "r01b_ubench_lines_per_instruction.txt); reduction lines: ncu lts__t_requests_srcunit_tex_op_red of the "
                      "shipped backward at configs[1] (profiles/traffic.json); live rows counted from this run's locations"}


def other_configs_table(msda, W, dev, iters=10):
    """BASELINE.json configs[2] (decoder, bf16, 50 / 550 queries, batch 16) and configs[4] (large-image shapes, batch 4,
    fp32 + bf16) timed on this GPU with the same kernels: parity-tested elsewhere, reported here so that the driver's
    record has them (N = 1 only; no roofline claim: these shapes are launch-latency or small-grid bound)."""
    rows = []
    wls = [W.config(2, num_queries=50), W.config(2, num_queries=550)] + W.sweep_config5(batch=4)
    for wl in wls:
        d = W.make_inputs(wl, device=dev)
        ab = W.algorithmic_bytes(wl)
        a5 = (d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"])

        def timeit(fn):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters

        f = timeit(lambda: torch.ops.msda.forward(*a5, 64))
        b = timeit(lambda: torch.ops.msda.backward(*a5, d["grad_out"], 64))
        rows.append({"workload": wl.name, "dtype": str(wl.dtype).replace("torch.", ""), "batch": wl.batch, "S": wl.S,
                     "queries": wl.Lq, "fwd_us": round(f * 1e3, 1), "bwd_us": round(b * 1e3, 1),
                     "GBps": round(ab["total"] / (f + b) / 1e6, 1),
                     "kernels": [msda._lib.describe("forward", wl.dtype, wl.batch, wl.heads, wl.head_dim, wl.L, wl.points, wl.Lq),
                                 msda._lib.describe("backward", wl.dtype, wl.batch, wl.heads, wl.head_dim, wl.L, wl.points, wl.Lq)]})
        del d, a5
        torch.cuda.empty_cache()
    return rows


def load_traffic(kernel):
    """dram bytes per launch from the committed ncu capture, if one has been summarised."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:  # noqa: BLE001
        return None


def time_reference_cuda(d, ab, args):
    """The reference's own CUDA kernels recompiled for sm_100a (oracle/_ref), same inputs, same
    event timing -- the on-GPU baseline of BASELINE.md 2b.  Reported beside, never part of, `value`."""
    try:
        from oracle import msda_oracle as O
        if not O.ref_cuda_available():
            return {"unavailable": "oracle/_ref/libmsda_ref_sm100.so not built"}
        a5 = (d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"])
        for _ in range(2):
            O.ref_cuda_forward(*a5); O.ref_cuda_backward(*a5, d["grad_out"])
        n = max(3, min(args.steps, 10))
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        fw = bw = 0.0
        for _ in range(n):
            e[0].record(); O.ref_cuda_forward(*a5); e[1].record(); O.ref_cuda_backward(*a5, d["grad_out"]); e[2].record()
            torch.cuda.synchronize()
            fw += e[0].elapsed_time(e[1]); bw += e[1].elapsed_time(e[2])
        fw, bw = fw / n, bw / n
        return {"fwd_ms": fw, "bwd_ms": bw, "GBps": ab["total"] / ((fw + bw) * 1e-3) / 1e9,
                "what": "reference ms_deform_im2col_cuda.cuh kernels, nvcc sm_100a, incl. their at::zeros-equivalent memsets"}
    except Exception as exc:  # noqa: BLE001
        return {"unavailable": repr(exc)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--loc-mode", default="model", choices=["model", "uniform"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--e2e-chunks", type=int, default=8)
    ap.add_argument("--e2e-images-per-chunk", type=int, default=2)
    ap.add_argument("--e2e-stages", type=int, default=3, help="device stages of the host-step pipeline (3..16)")
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--tune", default="", help="A/B only: comma-separated key=value for msda_set_tuning")
    ap.add_argument("--no-train-step", action="store_true", help="skip the MonoDETR training-step section (configs[3])")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the configs[2] / configs[4] timing table")
    ap.add_argument("--sustain-steps", type=int, default=100)
    ap.add_argument("--train-steps", type=int, default=20)
    ap.add_argument("--train-warmup", type=int, default=6)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        W = load_workloads()                        # never imports the package: the product library stays unloaded
        run_reference_arm(args, W.config(1, batch=args.batch, loc_mode=args.loc_mode))
    else:
        from monosowa_b200 import workloads as W
        run_ours(args, W.config(1, batch=args.batch, loc_mode=args.loc_mode))


if __name__ == "__main__":
    main()
