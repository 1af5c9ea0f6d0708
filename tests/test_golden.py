"""CPU: the oracle reproduces the committed golden vectors (tests/golden/*.npz, made by make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _z(name):
    return np.load(os.path.join(GOLDEN, name))


def test_integer_pipelines_bit_exact(oracle):
    z = _z("blur_small.npz")
    assert np.array_equal(oracle.blur(z["input"]), z["output"])
    z = _z("stencil_chain_small.npz")
    assert np.array_equal(oracle.stencil_chain(z["input"]), z["output"])
    z = _z("local_laplacian_small.npz")
    assert np.array_equal(oracle.local_laplacian(z["input"], int(z["levels"]), float(z["alpha"]), float(z["beta"])), z["output"])
    z = _z("camera_pipe_small.npz")
    got = oracle.camera_pipe(z["input"], z["m3200"], z["m7000"], 3700.0, 2.0, 50.0, 1.0, 25, 1023, z["output"].shape)
    assert np.array_equal(got, z["output"])


def test_float_pipelines(oracle):
    z = _z("bilateral_grid_small.npz")
    assert np.allclose(oracle.bilateral_grid(z["input"], float(z["r_sigma"])), z["output"], rtol=1e-6, atol=0)
    z = _z("nl_means_small.npz")
    got = oracle.nl_means(z["input"], int(z["patch"]), int(z["search"]), float(z["sigma"]))
    assert np.allclose(got, z["output"], rtol=1e-6, atol=0)
