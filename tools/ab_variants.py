"""A/B timing of the stencil_chain and nl_means kernel variants (hooks halide_b200_*_variant: 1 = first kernel,
2 = register-window kernel) at their bench_all sizes, and a bit-for-bit comparison of the two variants' outputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import halide_b200
from halide_b200 import HalideBuffer, filters
l = halide_b200.load_library()
g = torch.Generator(device="cuda"); g.manual_seed(3)


def timeit(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


# stencil_chain 1536 x 2560 u16
t_in = torch.randint(-32768, 32768, (2560, 1536), dtype=torch.int16, device="cuda", generator=g).view(torch.uint16)
outs = {}
for v in (1, 2, 1, 2):
    l.halide_b200_stencil_chain_variant(v)
    t_out = torch.zeros_like(t_in)
    bi, bo = HalideBuffer.from_torch(t_in), HalideBuffer.from_torch(t_out)
    us = timeit(lambda: filters.stencil_chain(bi, bo), 10)
    outs[v] = t_out.view(torch.int16).clone()
    print(f"stencil_chain 1536x2560 variant {v}: {us:.1f} us/call", flush=True)
l.halide_b200_stencil_chain_variant(0)
print("stencil_chain variants bit-identical:", bool(torch.equal(outs[1], outs[2])), flush=True)

# nl_means 3840 x 2160 x 3 f32, patch 3 search 7
f_in = torch.rand((3, 2160, 3840), dtype=torch.float32, device="cuda", generator=g)
outs = {}
for v in (1, 2, 1, 2):
    l.halide_b200_nl_means_variant(v)
    f_out = torch.zeros_like(f_in)
    bi, bo = HalideBuffer.from_torch(f_in), HalideBuffer.from_torch(f_out)
    us = timeit(lambda: filters.nl_means(bi, 3, 7, 0.12, bo), 3)
    outs[v] = f_out.clone()
    print(f"nl_means 3840x2160x3 variant {v}: {us:.1f} us/call", flush=True)
l.halide_b200_nl_means_variant(0)
print("nl_means variants bit-identical:", bool(torch.equal(outs[1], outs[2])), flush=True)
