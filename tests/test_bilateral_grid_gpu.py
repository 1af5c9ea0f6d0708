"""GPU parity tests for bilateral_grid through the C ABI.  Float pipeline: tolerance 1e-4 relative
(BASELINE.json north_star), checked against oracle/oracle_bilateral_grid.cpp."""
import numpy as np
import pytest

from util import f32_frame

pytestmark = pytest.mark.gpu
RTOL = 1e-4  # the tolerance north_star states for the float pipelines


def run(hb, inp, r_sigma, out_shape=None, in_mins=None, out_mins=None):
    out = np.zeros(inp.shape if out_shape is None else out_shape, np.float32)
    bi = hb.HalideBuffer.from_numpy(inp, in_mins)
    bo = hb.HalideBuffer.from_numpy(out, out_mins, host_dirty=False)
    hb.filters.bilateral_grid(bi, r_sigma, bo)
    bo.copy_to_host()
    return out


def close(got, want):
    err = np.abs(got - want) / np.maximum(np.abs(want), 1e-6)
    assert np.isfinite(got).all()
    assert err.max() <= RTOL, f"max rel err {err.max()} at {np.unravel_index(err.argmax(), err.shape)}"


@pytest.mark.parametrize("h,w", [(1, 1), (8, 8), (9, 23), (64, 128), (100, 257), (240, 519)])
def test_matches_oracle(hb, oracle, h, w):
    inp = f32_frame((h, w), h + w)
    close(run(hb, inp, 0.1), oracle.bilateral_grid(inp, 0.1))


@pytest.mark.parametrize("r_sigma", [0.05, 0.1, 0.25, 1.0])
def test_r_sigma_sweep(hb, oracle, r_sigma):
    inp = f32_frame((96, 160), 3)
    close(run(hb, inp, r_sigma), oracle.bilateral_grid(inp, r_sigma))


def test_out_of_range_values_are_clamped(hb, oracle):
    inp = (f32_frame((64, 96), 5) * 3.0 - 1.0).astype(np.float32)  # values in [-1, 2]
    close(run(hb, inp, 0.1), oracle.bilateral_grid(inp, 0.1))


def test_crop_with_offsets(hb, oracle):
    inp = f32_frame((120, 150), 8)
    kw = dict(out_shape=(70, 90), in_mins=(-13, 5), out_mins=(3, 22))
    close(run(hb, inp, 0.1, **kw), oracle.bilateral_grid(inp, 0.1, **kw))


def test_constant_frame_is_fixed_point(hb):
    """8K config size property: a constant frame is reproduced (weights cancel in the ratio)."""
    inp = np.full((4320, 7680), 0.37, np.float32)
    got = run(hb, inp, 0.1)
    assert np.allclose(got, 0.37, rtol=1e-5)


def test_8k_random_frame_regions(hb, oracle):
    """Config 3 size (7680 x 4320) on a RANDOM frame: regions of the full-size result against the oracle run on the crop
    that determines them (grid cells are 8 px, the blurs reach 2 cells, the slice 1: a 64 px margin is ample; crops keep
    their frame coordinates as mins because cell boundaries are absolute)."""
    h, w = 4320, 7680
    inp = f32_frame((h, w), 21)
    got = run(hb, inp, 0.1)
    for (y0, x0) in [(0, 0), (h - 96, w - 96), (2000, 3333), (4096 - 48, 4096 - 48)]:
        n, m = 96, 64
        ya, yb, xa, xb = max(0, y0 - m), min(h, y0 + n + m), max(0, x0 - m), min(w, x0 + n + m)
        crop = np.ascontiguousarray(inp[ya:yb, xa:xb])
        want = oracle.bilateral_grid(crop, 0.1, in_mins=(xa, ya), out_mins=(xa, ya))
        close(got[y0:y0 + n, x0:x0 + n], want[y0 - ya:y0 - ya + n, x0 - xa:x0 - xa + n])
