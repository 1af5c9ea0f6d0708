"""Generates tests/golden/*.npz by running the CPU oracle once on seeded inputs.

The reference holds no golden outputs for these pipelines and cannot be built in
this image, so the vectors pin OUR restatement: they detect the oracle and the kernels drifting
together (tests/test_golden.py checks the oracle on CPU, tests/test_golden_gpu.py the kernels).
Re-run only when the oracle's definition is deliberately changed:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle  # noqa: E402

M3200 = np.array([[1.6697, -0.2693, -0.4004, -42.4346], [-0.3576, 1.0615, 1.5949, -37.1158],
                  [-0.2175, -1.8751, 6.9640, -26.6970]], np.float32)
M7000 = np.array([[2.2997, -0.4478, 0.1706, -39.0923], [-0.3826, 1.5906, -0.2080, -25.4311],
                  [-0.0888, -0.7344, 2.2832, -20.0826]], np.float32)


def save(name, **kw):
    np.savez_compressed(os.path.join(HERE, name), **kw)


def main():
    rng = np.random.default_rng(20260922)
    img = rng.integers(0, 65536, (3, 48, 80), dtype=np.uint16)
    levels, alpha, beta = 8, np.float32(1.0 / 7.0), np.float32(1.0)
    save("local_laplacian_small.npz", input=img, output=pyoracle.local_laplacian(img, levels, float(alpha), float(beta)),
         levels=levels, alpha=alpha, beta=beta)
    b_in = rng.integers(0, 65536, (34, 72), dtype=np.uint16)
    save("blur_small.npz", input=b_in, output=pyoracle.blur(b_in))
    s_in = rng.integers(0, 65536, (40, 56), dtype=np.uint16)
    save("stencil_chain_small.npz", input=s_in, output=pyoracle.stencil_chain(s_in))
    g_in = rng.random((48, 72), dtype=np.float32)
    save("bilateral_grid_small.npz", input=g_in, r_sigma=np.float32(0.1), output=pyoracle.bilateral_grid(g_in, 0.1))
    n_in = rng.random((3, 24, 40), dtype=np.float32)
    save("nl_means_small.npz", input=n_in, patch=3, search=7, sigma=np.float32(0.12),
         output=pyoracle.nl_means(n_in, 3, 7, 0.12))
    raw = rng.integers(0, 1024, (88, 96), dtype=np.uint16)
    shape = (3, ((88 - 24) // 32) * 32, ((96 - 32) // 32) * 32)
    save("camera_pipe_small.npz", input=raw, m3200=M3200, m7000=M7000,
         output=pyoracle.camera_pipe(raw, M3200, M7000, 3700.0, 2.0, 50.0, 1.0, 25, 1023, shape))
    print("wrote golden vectors")


if __name__ == "__main__":
    main()
