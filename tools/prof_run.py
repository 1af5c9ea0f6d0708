"""Minimal driver for ncu captures: runs one filter a few times on device-resident synthetic frames.
    python tools/prof_run.py local_laplacian 3840 2160 3     (name W H steps)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from halide_b200 import HalideBuffer, filters  # noqa: E402


def main():
    name = sys.argv[1]
    W, H, steps = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    if name == "local_laplacian":
        t = torch.randint(-32768, 32768, (3, H, W), dtype=torch.int16, device=dev, generator=g).view(torch.uint16)
        o = torch.zeros_like(t)
        bi, bo = HalideBuffer.from_torch(t), HalideBuffer.from_torch(o)
        run = lambda: filters.local_laplacian(bi, 8, 1.0 / 7.0, 1.0, bo)
    elif name == "blur":
        t = torch.randint(-32768, 32768, (H + 2, W + 2), dtype=torch.int16, device=dev, generator=g).view(torch.uint16)
        o = torch.zeros((H, W), dtype=torch.uint16, device=dev)
        bi, bo = HalideBuffer.from_torch(t), HalideBuffer.from_torch(o)
        run = lambda: filters.halide_blur(bi, bo)
    elif name == "bilateral_grid":
        t = torch.rand((H, W), dtype=torch.float32, device=dev, generator=g)
        o = torch.zeros_like(t)
        bi, bo = HalideBuffer.from_torch(t), HalideBuffer.from_torch(o)
        run = lambda: filters.bilateral_grid(bi, 0.1, bo)
    elif name == "stencil_chain":
        t = torch.randint(-32768, 32768, (H, W), dtype=torch.int16, device=dev, generator=g).view(torch.uint16)
        o = torch.zeros_like(t)
        bi, bo = HalideBuffer.from_torch(t), HalideBuffer.from_torch(o)
        run = lambda: filters.stencil_chain(bi, bo)
    elif name == "nl_means":
        t = torch.rand((3, H, W), dtype=torch.float32, device=dev, generator=g)
        o = torch.zeros_like(t)
        bi, bo = HalideBuffer.from_torch(t), HalideBuffer.from_torch(o)
        run = lambda: filters.nl_means(bi, 3, 7, 0.12, bo)
    elif name == "conv_layer":
        ti = torch.rand((5, 82, 102, 128), dtype=torch.float32, device=dev, generator=g)
        tf = torch.rand((128, 3, 3, 128), dtype=torch.float32, device=dev, generator=g)
        tb = torch.rand((128,), dtype=torch.float32, device=dev, generator=g)
        o = torch.zeros((5, 80, 100, 128), dtype=torch.float32, device=dev)
        bi, bf, bb, bo = (HalideBuffer.from_torch(t) for t in (ti, tf, tb, o))
        run = lambda: filters.conv_layer(bi, bf, bb, bo)
    elif name == "camera_pipe":
        import numpy as np
        t = torch.randint(0, 1024, (H + 48, W + 64), dtype=torch.int16, device=dev, generator=g).view(torch.uint16)
        m32 = torch.tensor([[1.6697, -0.2693, -0.4004, -42.4346], [-0.3576, 1.0615, 1.5949, -37.1158],
                            [-0.2175, -1.8751, 6.9640, -26.6970]], dtype=torch.float32, device=dev)
        m70 = torch.tensor([[2.2997, -0.4478, 0.1706, -39.0923], [-0.3826, 1.5906, -0.2080, -25.4311],
                            [-0.0888, -0.7344, 2.2832, -20.0826]], dtype=torch.float32, device=dev)
        o = torch.zeros((3, H, W), dtype=torch.uint8, device=dev)
        bi, b32, b70, bo = (HalideBuffer.from_torch(x) for x in (t, m32, m70, o))
        run = lambda: filters.camera_pipe(bi, b32, b70, 3700.0, 2.0, 50.0, 1.0, 25, 1023, bo)
    else:
        raise SystemExit(f"unknown filter {name}")
    for _ in range(steps):
        run()
    torch.cuda.synchronize()
    # event-timed loop (printed for convenience; never a bench number when run under a profiler)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name} {W}x{H}: {e0.elapsed_time(e1) / steps * 1e3:.1f} us/step")


if __name__ == "__main__":
    main()
