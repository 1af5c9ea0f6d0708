"""A/B of the level-1 kernels on one GPU: ll_down_strip_kernel<8,true> (default) against the experimental
ll_level1_pair_kernel (halide_b200_ll_force_generic(32)) on a 4K frame: bit-compare the outputs, then per-kernel times
from the library's event-bracketed profile.    python tools/level1_ab.py [W H]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import halide_b200  # noqa: E402
from halide_b200 import HalideBuffer, filters, lib as hlib  # noqa: E402


def main():
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 3840
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 2160
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    t = torch.randint(-32768, 32768, (3, H, W), dtype=torch.int16, device=dev, generator=g).view(torch.uint16)
    outs, reports = {}, {}
    for mode in (0, 32):
        halide_b200.capi.halide_b200_ll_force_generic(mode)
        o = torch.zeros_like(t)
        bi, bo = HalideBuffer.from_torch(t), HalideBuffer.from_torch(o)
        for _ in range(3):
            filters.local_laplacian(bi, 8, 1.0 / 7.0, 1.0, bo)
        torch.cuda.synchronize()
        hlib.profile(True)
        hlib.profile_reset()
        for _ in range(20):
            filters.local_laplacian(bi, 8, 1.0 / 7.0, 1.0, bo)
        torch.cuda.synchronize()
        reports[mode] = {k: round(ms / c * 1e3, 1) for k, (c, ms) in hlib.profile_report().items()}
        hlib.profile(False)
        outs[mode] = o
    halide_b200.capi.halide_b200_ll_force_generic(0)
    diff = int((outs[0].view(torch.int16) != outs[32].view(torch.int16)).sum().item())
    print("LEVEL1_AB frame=%dx%d mismatching_samples=%d" % (W, H, diff))
    print("default us/launch:", reports[0])
    print("pair    us/launch:", reports[32])


if __name__ == "__main__":
    main()
