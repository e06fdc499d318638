"""Extended randomised parity run (beyond tests/test_msda_gpu.py::test_fuzz_random_shapes): random pyramids, query sets
(pixel pyramids and arbitrary), dtypes, kernel families and the fused op with 2- / 6-dim reference points, each against
the fp64 C oracle.   python tools/fuzz_gpu.py [--cases 300] [--seed 1]   (MSDA_AB=1 adds the tile kernels)"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import monosowa_b200 as msda  # noqa: E402
from monosowa_b200.ops.functions import MSDeformAttnFusedFunction  # noqa: E402
from monosowa_b200.ops.modules.ms_deform_attn import sampling_locations_from_reference  # noqa: E402
from oracle import msda_oracle as O  # noqa: E402

TOL = {torch.float64: dict(fwd=1e-12, gv=1e-12, ga=1e-12, gl=1e-11), torch.float32: dict(fwd=1e-5, gv=1e-4, ga=1e-4, gl=1e-4),
       torch.bfloat16: dict(fwd=4e-3, gv=1e-2, ga=1e-4, gl=1e-4)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(a.seed)
    r = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    fams = [(-1, -1), (11, 11), (11, 21), (99, 99)] + ([(12, 20)] if msda._lib.has_ab_flavours() else [])
    bad = 0
    for case in range(a.cases):
        L = r(1, 4)
        shapes = [(r(1, 30), r(1, 40)) for _ in range(L)]
        D = [16, 32, 64, 32, 32, 24][r(0, 5)]
        dtype = [torch.float32, torch.float32, torch.bfloat16, torch.float64][r(0, 3)]
        N, M, P = r(1, 3), [1, 2, 3, 8][r(0, 3)], r(1, 5)
        sh = torch.tensor(shapes)
        lsi = torch.cat((sh.new_zeros(1), sh.prod(1).cumsum(0)[:-1]))
        S = int(sh.prod(1).sum())
        pyramid = r(0, 1) == 1
        Lq = S if pyramid else r(1, 300)
        fam = fams[r(0, len(fams) - 1)]
        fused = r(0, 2) == 0 and dtype != torch.float64 and D in (16, 32, 64) and L * P <= D and fam[0] != 99   # (99 forces the generic kernels: no fused flavour)
        ref_dim = [2, 6][r(0, 1)]
        ct = torch.float64 if dtype == torch.float64 else torch.float32
        value = torch.randn(N, S, M, D, generator=g, dtype=torch.float64).to(dtype)
        grad_out = torch.randn(N, Lq, M * D, generator=g, dtype=torch.float64).to(dtype)
        logits = torch.randn(N, Lq, M, L * P, generator=g, dtype=torch.float64).to(ct)
        if fused:
            ref = (torch.rand(N, Lq, L, 2, generator=g) * 1.2 - 0.1)
            if ref_dim == 6:
                ref = torch.cat([ref, torch.rand(N, Lq, L, 4, generator=g) * 0.3 + 0.02], -1)
            offs = torch.randn(N, Lq, M, L, P, 2, generator=g) * 3.0
            loc = sampling_locations_from_reference(ref, offs, sh, P).to(ct)
        else:
            loc = (torch.rand(N, Lq, M, L, P, 2, generator=g, dtype=torch.float64) * 1.5 - 0.25).to(ct)
        attn = torch.softmax(logits, -1).view(N, Lq, M, L, P)
        msda._lib.set_tuning("fwd_variant", fam[0]); msda._lib.set_tuning("bwd_variant", fam[1])
        try:
            v = value.to(dev).requires_grad_(True)
            if fused:
                o = offs.to(dev).requires_grad_(True); lg = logits.to(dev).requires_grad_(True)
                out = MSDeformAttnFusedFunction.apply(v, sh.to(dev), lsi.to(dev), ref.to(dev), o, lg)
                out.backward(grad_out.to(dev))
                loc64, attn64 = loc.double(), attn.double()
                got = dict(fwd=out.detach().cpu(), gv=v.grad.cpu())
            else:
                l = loc.to(dev).requires_grad_(True); at = attn.to(dev).requires_grad_(True)
                out = msda.MSDeformAttnFunction.apply(v, sh.to(dev), lsi.to(dev), l, at, 64)
                out.backward(grad_out.to(dev))
                loc64, attn64 = loc.double(), attn.double()
                got = dict(fwd=out.detach().cpu(), gv=v.grad.cpu(), ga=at.grad.cpu(), gl=l.grad.cpu())
        finally:
            msda._lib.set_tuning("fwd_variant", -1); msda._lib.set_tuning("bwd_variant", -1)
        args = (value.double(), sh, lsi, loc64, attn64)
        want = dict(fwd=O.forward_c(*args))
        want["gv"], want["gl"], want["ga"] = O.backward_c(*args, grad_out.double())
        keep = ~O.pixel_boundary_mask(loc, sh, eps_px=1e-4)
        errs = {}
        for k in got:
            errs[k] = O.rel_l2(got[k][keep], want[k][keep]) if k == "gl" else O.rel_l2(got[k], want[k])
        # the fused op forms its locations on the device in fp32: compare within the fp32 location noise
        tol = dict(TOL[dtype])
        if fused:
            tol["fwd"] = max(tol["fwd"], 2e-5); tol["gv"] = max(tol["gv"], 1e-4)
        fail = {k: e for k, e in errs.items() if not e <= tol[k]}
        if fail:
            bad += 1
            print(f"case {case}: FAIL {fail} shapes={shapes} N={N} M={M} D={D} P={P} Lq={Lq} {dtype} fam={fam} fused={fused} ref_dim={ref_dim}", flush=True)
    print(f"fuzz: {a.cases} cases, {bad} failures")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
