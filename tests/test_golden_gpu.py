"""GPU: the C-ABI filters reproduce the committed golden vectors (independent of building the oracle)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _z(name):
    return np.load(os.path.join(GOLDEN, name))


def _call(hb, fn, inputs, out, *scalars_between):
    bufs = [hb.HalideBuffer.from_numpy(np.ascontiguousarray(a)) for a in inputs]
    bo = hb.HalideBuffer.from_numpy(out, host_dirty=False)
    fn(*bufs, *scalars_between, bo)
    bo.copy_to_host()
    return out


def test_blur(hb):
    z = _z("blur_small.npz")
    assert np.array_equal(_call(hb, hb.filters.halide_blur, [z["input"]], np.zeros_like(z["output"])), z["output"])


def test_stencil_chain(hb):
    z = _z("stencil_chain_small.npz")
    assert np.array_equal(_call(hb, hb.filters.stencil_chain, [z["input"]], np.zeros_like(z["output"])), z["output"])


def test_local_laplacian(hb):
    z = _z("local_laplacian_small.npz")
    got = _call(hb, hb.filters.local_laplacian, [z["input"]], np.zeros_like(z["output"]), int(z["levels"]), float(z["alpha"]),
                float(z["beta"]))
    assert np.array_equal(got, z["output"])


def test_camera_pipe(hb):
    z = _z("camera_pipe_small.npz")
    got = _call(hb, hb.filters.camera_pipe, [z["input"], z["m3200"], z["m7000"]], np.zeros_like(z["output"]), 3700.0, 2.0, 50.0,
                1.0, 25, 1023)
    assert np.array_equal(got, z["output"])


def test_bilateral_grid(hb):
    z = _z("bilateral_grid_small.npz")
    got = _call(hb, hb.filters.bilateral_grid, [z["input"]], np.zeros_like(z["output"]), float(z["r_sigma"]))
    assert np.max(np.abs(got - z["output"]) / np.maximum(np.abs(z["output"]), 1e-6)) <= 1e-4


def test_nl_means(hb):
    z = _z("nl_means_small.npz")
    got = _call(hb, hb.filters.nl_means, [z["input"]], np.zeros_like(z["output"]), int(z["patch"]), int(z["search"]),
                float(z["sigma"]))
    assert np.max(np.abs(got - z["output"]) / np.maximum(np.abs(z["output"]), 1e-3)) <= 1e-4
