"""Generates tests/golden/*.npz by running the CPU oracle once on seeded inputs.

The reference holds no golden outputs for these pipelines and cannot be built in
this image, so the vectors pin OUR restatement: they detect the oracle and the kernels drifting
together.  Re-run only when the oracle's definition is deliberately changed:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle  # noqa: E402


def main():
    rng = np.random.default_rng(20260922)
    img = rng.integers(0, 65536, (3, 48, 80), dtype=np.uint16)
    levels, alpha, beta = 8, np.float32(1.0 / 7.0), np.float32(1.0)
    out = pyoracle.local_laplacian(img, levels, float(alpha), float(beta))
    np.savez_compressed(os.path.join(HERE, "local_laplacian_small.npz"), input=img, output=out, levels=levels,
                        alpha=alpha, beta=beta)
    b_in = (rng.integers(0, 65536, (34, 72), dtype=np.uint16))
    np.savez_compressed(os.path.join(HERE, "blur_small.npz"), input=b_in, output=pyoracle.blur(b_in))
    print("wrote golden vectors")


if __name__ == "__main__":
    main()
