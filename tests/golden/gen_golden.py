"""Generate tests/golden/*.npz by running the REFERENCE's own code in the build container.

    python tests/golden/gen_golden.py        # needs /root/reference; rewrites the fixtures

What is imported from /root/reference (never copied):
  * ``ms_deform_attn_core_pytorch``  MonoDETR/lib/models/monodetr/ops/functions/ms_deform_attn_func.py:41-61
    -- the function the reference's own ops/test.py pins its CUDA kernels against;
  * ``MSDeformAttn``                 MonoDETR/lib/models/monodetr/ops/modules/ms_deform_attn.py:69-162
    -- run on CPU with ``MSDeformAttnFunction.apply`` routed to the function above, to pin the
    module-level arithmetic (softmax over L*P, 2-dim and 6-dim reference points).
Import shims: the compiled extension ``MultiScaleDeformableAttention`` is
stubbed in ``sys.modules`` (func.py:18 imports it at module top), and two torch-version
bugs at ms_deform_attn.py:34/:55 are satisfied with aliases.  No reference file is edited.

Cases follow the reference's ops/test.py (seed 3, levels (6,4),(3,2), N=1, M=2, Lq=2, L=2,
P=2, value=rand*0.01, loc=rand, normalised weights; D in its gradcheck list) and add the
shapes this repo cares about (D=32, L=P=4, out-of-range locations, a KITTI-pyramid slice).
Each case stores the inputs (fp64), the reference output and the autograd gradients for a
stored grad_output, in fp64 and (forward only) fp32.
"""
import os
import sys
import types

import numpy as np
import torch

REF_OPS = "/root/reference/MonoDETR/lib/models/monodetr/ops"
OUT_DIR = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    import torch.nn.modules.linear as _lin
    if not hasattr(_lin, "_LinearWithBias"):
        _lin._LinearWithBias = _lin.NonDynamicallyQuantizableLinear
    if "torch._overrides" not in sys.modules:
        import torch.overrides as _ov
        fake = types.ModuleType("torch._overrides")
        fake.has_torch_function = _ov.has_torch_function
        fake.handle_torch_function = _ov.handle_torch_function
        sys.modules["torch._overrides"] = fake
    # import as a package "refops" so that the module's relative import (..functions) works
    import importlib.util
    def load_pkg(name, path):
        spec = importlib.util.spec_from_file_location(name, os.path.join(path, "__init__.py"),
                                                      submodule_search_locations=[path])
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        return spec, mod
    root = types.ModuleType("refops"); root.__path__ = [REF_OPS]; sys.modules["refops"] = root
    spec_f, functions = load_pkg("refops.functions", os.path.join(REF_OPS, "functions"))
    spec_f.loader.exec_module(functions)
    spec_m, modules = load_pkg("refops.modules", os.path.join(REF_OPS, "modules"))
    spec_m.loader.exec_module(modules)
    from refops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch, MSDeformAttnFunction
    return ms_deform_attn_core_pytorch, MSDeformAttnFunction, modules.MSDeformAttn


def level_index(shapes):
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def op_case(core, name, seed, shapes, N, M, D, Lq, P, loc_mode="unit", value_scale=0.01):
    g = torch.Generator().manual_seed(seed)
    shapes = torch.as_tensor(shapes, dtype=torch.long)
    L = shapes.shape[0]
    S = int(shapes.prod(1).sum())
    value = torch.rand(N, S, M, D, generator=g, dtype=torch.float64) * value_scale
    if loc_mode == "unit":            # ops/test.py:33
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g, dtype=torch.float64)
    elif loc_mode == "oob":           # what the hooked model produces: ~15 % outside [0,1]
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g, dtype=torch.float64) * 2.32 - 0.66
    elif loc_mode == "edges":         # exact borders / pixel centres / just outside
        base = torch.tensor([-0.5, -1e-3, 0.0, 1e-3, 0.25, 0.5, 0.999, 1.0, 1.001, 1.5], dtype=torch.float64)
        idx = torch.randint(0, len(base), (N, Lq, M, L, P, 2), generator=g)
        loc = base[idx]
    else:
        raise ValueError(loc_mode)
    attn = torch.rand(N, Lq, M, L, P, generator=g, dtype=torch.float64) + 1e-5
    attn = attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)           # ops/test.py:35
    grad_out = torch.randn(N, Lq, M * D, generator=g, dtype=torch.float64)

    v, l, a = (t.clone().requires_grad_(True) for t in (value, loc, attn))
    out64 = core(v, shapes, l, a)
    out64.backward(grad_out)
    out32 = core(value.float(), shapes, loc.float(), attn.float())
    np.savez_compressed(
        os.path.join(OUT_DIR, f"op_{name}.npz"),
        shapes=shapes.numpy(), level_start_index=level_index(shapes).numpy(),
        value=value.numpy(), loc=loc.numpy(), attn=attn.numpy(), grad_out=grad_out.numpy(),
        out64=out64.detach().numpy(), out32=out32.numpy(),
        grad_value=v.grad.numpy(), grad_loc=l.grad.numpy(), grad_attn=a.grad.numpy())
    print(f"op_{name}: N={N} S={S} M={M} D={D} L={L} Lq={Lq} P={P} |out|={out64.abs().max():.3e}")


def module_case(core, Function, MSDeformAttn, name, seed, ref_dim, shapes, N, Lq, d_model=64, heads=4, P=2):
    torch.manual_seed(seed)
    shapes = torch.as_tensor(shapes, dtype=torch.long)
    L = shapes.shape[0]
    S = int(shapes.prod(1).sum())
    mod = MSDeformAttn(d_model=d_model, n_levels=L, n_heads=heads, n_points=P).double()
    # the default init zeroes two of the Linears; perturb so every path carries signal
    with torch.no_grad():
        for prm in mod.parameters():
            prm.add_(torch.randn_like(prm) * 0.05)
    query = torch.randn(N, Lq, d_model, dtype=torch.float64)
    src = torch.randn(N, S, d_model, dtype=torch.float64)
    ref = torch.rand(N, Lq, L, ref_dim, dtype=torch.float64)
    if ref_dim == 6:
        ref[..., 2:] *= 0.3
    mask = torch.zeros(N, S, dtype=torch.bool)
    mask[:, -3:] = True
    # route the autograd Function to the reference's own PyTorch core (the ext is a stub here)
    orig_apply = Function.apply
    Function.apply = staticmethod(lambda v, sh, lsi, loc, aw, step: core(v, sh, loc, aw))
    try:
        out = mod(query, ref, src, shapes, level_index(shapes), mask)
    finally:
        Function.apply = orig_apply
    state = {k: v.detach().numpy() for k, v in mod.state_dict().items()}
    np.savez_compressed(
        os.path.join(OUT_DIR, f"module_{name}.npz"),
        shapes=shapes.numpy(), level_start_index=level_index(shapes).numpy(),
        query=query.numpy(), src=src.numpy(), ref=ref.numpy(), mask=mask.numpy(),
        out=out.detach().numpy(), d_model=d_model, heads=heads, points=P,
        **{"state__" + k: v for k, v in state.items()})
    print(f"module_{name}: ref_dim={ref_dim} out {tuple(out.shape)}")


def main():
    core, Function, MSDeformAttn = import_reference()
    tiny = [(6, 4), (3, 2)]                                   # ops/test.py:24
    for d in (2, 4, 30, 32, 64, 71, 1025):                    # ops/test.py:21,85 (+ D=2 of :21)
        op_case(core, f"testpy_D{d}", 3, tiny, N=1, M=2, D=d, Lq=2, P=2)
    op_case(core, "d32_l4p4_unit", 11, [(8, 20), (4, 10), (2, 5), (1, 3)], N=2, M=4, D=32, Lq=37, P=4)
    op_case(core, "d32_l4p4_oob", 12, [(8, 20), (4, 10), (2, 5), (1, 3)], N=2, M=4, D=32, Lq=37, P=4,
            loc_mode="oob", value_scale=1.0)
    op_case(core, "d32_edges", 13, [(5, 7), (3, 4)], N=1, M=3, D=32, Lq=29, P=3, loc_mode="edges", value_scale=1.0)
    op_case(core, "d16_l3p1_oob", 14, [(9, 11), (5, 6), (1, 1)], N=3, M=2, D=16, Lq=5, P=1, loc_mode="oob", value_scale=1.0)
    op_case(core, "d8_oob", 15, [(4, 4)], N=1, M=1, D=8, Lq=64, P=8, loc_mode="oob", value_scale=1.0)
    module_case(core, Function, MSDeformAttn, "ref2", 21, 2, [(6, 8), (3, 4)], N=2, Lq=7)
    module_case(core, Function, MSDeformAttn, "ref6", 22, 6, [(6, 8), (3, 4)], N=2, Lq=7)


if __name__ == "__main__":
    main()
