"""GPU parity tests for local_laplacian through the C ABI: bit-exact uint16 against the oracle
(oracle/oracle_local_laplacian.cpp, a restatement of apps/local_laplacian/local_laplacian_generator.cpp)."""
import os

import numpy as np
import pytest

from util import run_local_laplacian, smooth_u16_frame, u16_frame

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _check(hb, oracle, img, levels, alpha, beta, **kw):
    got = run_local_laplacian(hb, img, levels, alpha, beta, **kw)
    want = oracle.local_laplacian(img, levels, alpha, beta, **kw)
    bad = np.argwhere(got != want)
    assert bad.size == 0, (f"{len(bad)} mismatching samples, first at (c,y,x)={bad[0]}: "
                           f"got {got[tuple(bad[0])]} want {want[tuple(bad[0])]}")


@pytest.mark.parametrize("h,w", [(1, 1), (2, 3), (7, 5), (16, 16), (33, 47), (64, 96), (100, 130), (255, 257)])
def test_small_and_ragged_frames(hb, oracle, h, w):
    _check(hb, oracle, u16_frame((3, h, w), h * 1000 + w), 8, 1.0 / 7.0, 1.0)


@pytest.mark.parametrize("seed", [0, 42, 1, 2])
def test_random_full_range(hb, oracle, seed):
    _check(hb, oracle, u16_frame((3, 192, 320), seed), 8, 1.0 / 7.0, 1.0)


@pytest.mark.parametrize("seed", [0, 1])
def test_smooth_content(hb, oracle, seed):
    _check(hb, oracle, smooth_u16_frame((3, 200, 264), seed), 8, 1.0 / 7.0, 1.0)


@pytest.mark.parametrize("levels,alpha,beta", [(2, 1.0, 1.0), (4, 0.5, 0.7), (8, 2.0 / 7.0, 1.5), (8, 0.0, 1.0),
                                               (16, 1.0 / 15.0, 0.3)])
def test_parameter_sweep(hb, oracle, levels, alpha, beta):
    _check(hb, oracle, u16_frame((3, 72, 104), levels), levels, alpha, beta)


def test_extreme_pixels(hb, oracle):
    img = u16_frame((3, 64, 80), 9)
    img[:, :8] = 0
    img[:, 8:16] = 65535
    img[0, 16:24] = 0
    img[1, 16:24] = 65535
    _check(hb, oracle, img, 8, 1.0 / 7.0, 1.0)


def test_output_crop_with_offsets(hb, oracle):
    """Output region strictly inside a larger input with non-zero mins: the pyramid is clamped at the
    INPUT's edges (repeat_edge uses the buffer's own bounds, src/BoundaryConditions.cpp:15-35)."""
    img = u16_frame((3, 90, 120), 4)
    _check(hb, oracle, img, 8, 1.0 / 7.0, 1.0, out_shape=(3, 50, 61), in_mins=(-7, 3, 0), out_mins=(10, 21, 0))
    _check(hb, oracle, img, 8, 1.0 / 7.0, 1.0, out_shape=(3, 90, 120), in_mins=(5, -4, 0), out_mins=(5, -4, 0))


def test_golden_fixture(hb):
    """Committed vectors (tests/golden/make_golden.py ran the oracle once): guards both the oracle and
    the kernels against drifting together."""
    z = np.load(os.path.join(GOLDEN, "local_laplacian_small.npz"))
    got = run_local_laplacian(hb, z["input"], int(z["levels"]), float(z["alpha"]), float(z["beta"]))
    assert np.array_equal(got, z["output"])


def test_idempotent_and_deterministic(hb):
    img = u16_frame((3, 300, 420), 77)
    a = run_local_laplacian(hb, img, 8, 1.0 / 7.0, 1.0)
    b = run_local_laplacian(hb, img, 8, 1.0 / 7.0, 1.0)
    assert np.array_equal(a, b)


def test_4k_tile_consistency(hb, oracle):
    """Config 2 size (3840x2160x3).  The oracle needs ~1 min for a full 4K frame, so the full-size
    check is by region: the top-left and bottom-right 256x256 corners of the 4K result must equal
    the oracle run on the enlarged crop that determines them (footprint of 8 levels < 2*256+64 px)."""
    h, w = 2160, 3840
    img = u16_frame((3, h, w), 0)
    got = run_local_laplacian(hb, img, 8, 1.0 / 7.0, 1.0)
    m = 1024  # margin: level-7 taps reach < 2^8 * 3 px; the crop keeps the true frame edge on two sides
    tl = oracle.local_laplacian(np.ascontiguousarray(img[:, :m, :m]), 8, 1.0 / 7.0, 1.0)
    assert np.array_equal(got[:, :160, :160], tl[:, :160, :160])
    # pyramid coordinates are absolute (parity of x,y matters in down/upsample): keep the crop's mins
    br = oracle.local_laplacian(np.ascontiguousarray(img[:, h - m:, w - m:]), 8, 1.0 / 7.0, 1.0,
                                in_mins=(w - m, h - m, 0), out_mins=(w - m, h - m, 0))
    assert np.array_equal(got[:, h - 160:, w - 160:], br[:, m - 160:, m - 160:])


def test_generic_kernels_agree_with_fast_path(hb, oracle):
    """levels == 8 takes the warp-strip kernels; any other `levels` (and this hook) takes the generic
    per-pixel kernels.  Both must equal the oracle."""
    img = u16_frame((3, 131, 203), 31)
    want = oracle.local_laplacian(img, 8, 1.0 / 7.0, 1.0)
    l = hb.load_library()
    try:
        l.halide_b200_ll_force_generic(7)
        got_generic = run_local_laplacian(hb, img, 8, 1.0 / 7.0, 1.0)
    finally:
        l.halide_b200_ll_force_generic(0)
    got_fast = run_local_laplacian(hb, img, 8, 1.0 / 7.0, 1.0)
    assert np.array_equal(got_generic, want) and np.array_equal(got_fast, want)


@pytest.mark.parametrize("shape", [(3, 130, 256), (3, 97, 198), (3, 64, 66)])
def test_final_kernel_simple_and_general_layout_paths(hb, oracle, shape):
    """Even-width 3-channel frames with 4-byte aligned rows take the final kernel's SIMPLE path (32-bit addressing,
    one aligned word per thread and channel); hook bit 16 forces the general-layout path on the same frame.
    Both must equal the oracle; a crop with odd column offsets must fall back to the general path by itself."""
    img = u16_frame(shape, 77)
    want = oracle.local_laplacian(img, 8, 1.0 / 7.0, 1.0)
    l = hb.load_library()
    got_simple = run_local_laplacian(hb, img, 8, 1.0 / 7.0, 1.0)
    try:
        l.halide_b200_ll_force_generic(16)
        got_general = run_local_laplacian(hb, img, 8, 1.0 / 7.0, 1.0)
    finally:
        l.halide_b200_ll_force_generic(0)
    assert np.array_equal(got_simple, want) and np.array_equal(got_general, want)
    c, h, w = shape
    out_shape = (3, h - 9, w - 12)  # even width, output columns start at an odd input column
    want_crop = oracle.local_laplacian(img, 8, 1.0 / 7.0, 1.0, out_shape=out_shape, in_mins=(0, 0, 0), out_mins=(5, 4, 0))
    got_crop = run_local_laplacian(hb, img, 8, 1.0 / 7.0, 1.0, out_shape=out_shape, in_mins=(0, 0, 0), out_mins=(5, 4, 0))
    assert np.array_equal(got_crop, want_crop)
    out_shape = (3, h - 8, w - 12)  # even offsets: SIMPLE path on a crop (row pointer no longer at the buffer start)
    want_crop = oracle.local_laplacian(img, 8, 1.0 / 7.0, 1.0, out_shape=out_shape, in_mins=(0, 0, 0), out_mins=(6, 4, 0))
    got_crop = run_local_laplacian(hb, img, 8, 1.0 / 7.0, 1.0, out_shape=out_shape, in_mins=(0, 0, 0), out_mins=(6, 4, 0))
    assert np.array_equal(got_crop, want_crop)


@pytest.mark.parametrize("shape", [(3, 131, 203), (3, 130, 256), (3, 64, 66), (3, 257, 1031)])
def test_pair_column_level1_kernel(hb, oracle, shape):
    """ll_level1_pair_kernel (hook bit 32; two source columns per lane): bit-exact, kept as an alternative to
    ll_down_strip_kernel<8, true> although it measured slightly slower (tools/level1_ab.py)."""
    img = u16_frame(shape, 91)
    want = oracle.local_laplacian(img, 8, 1.0 / 7.0, 1.0)
    l = hb.load_library()
    try:
        l.halide_b200_ll_force_generic(32)
        got = run_local_laplacian(hb, img, 8, 1.0 / 7.0, 1.0)
    finally:
        l.halide_b200_ll_force_generic(0)
    assert np.array_equal(got, want)
