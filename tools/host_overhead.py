"""Host-side cost of one op call at a launch-bound shape (decoder, 50 queries): where do the microseconds go?"""
import ctypes, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import monosowa_b200 as msda
from monosowa_b200 import workloads as W
from monosowa_b200.ops.functions import ms_deform_attn_func as F

dev = torch.device("cuda:0")
d = W.make_inputs(W.config(2, num_queries=50, dtype=torch.float32), device=dev)
a5 = (d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"])

def bench(fn, n=2000):
    for _ in range(200): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize()
    return (t1 - t0) / n * 1e6

print("torch.ops.msda.forward      %.1f us" % bench(lambda: torch.ops.msda.forward(*a5, 64)))
print("  _forward_cuda direct      %.1f us" % bench(lambda: F._forward_cuda(*a5, 64)))
print("  _check                    %.1f us" % bench(lambda: F._check(*a5, 64)))
print("  torch.empty               %.1f us" % bench(lambda: torch.empty((16, 50, 256), dtype=torch.float32, device=dev)))
def raw():
    with F._on_device(dev) as st:
        pass
print("  _on_device ctx            %.1f us" % bench(raw))
out = torch.empty((16, 50, 256), device=dev)
fn = msda._lib.lib.msda_forward_f32
p = [ctypes.c_void_p(t.data_ptr()) for t in (*a5, out)]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
print("  ctypes call (launch)      %.1f us" % bench(lambda: fn(*p, 16, 10200, 8, 32, 4, 50, 4, st)))
print("  6x data_ptr+c_void_p      %.1f us" % bench(lambda: [ctypes.c_void_p(t.data_ptr()) for t in (*a5, out)]))
print("MSDeformAttnFunction.apply  %.1f us" % bench(lambda: msda.MSDeformAttnFunction.apply(*a5, 64)))
print("torch.ops.msda.backward     %.1f us" % bench(lambda: torch.ops.msda.backward(*a5, d["grad_out"], 64)))
