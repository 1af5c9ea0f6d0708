"""GPU parity tests for camera_pipe through the C ABI: bit-exact uint8 against oracle/oracle_camera_pipe.cpp."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

M3200 = np.array([[1.6697, -0.2693, -0.4004, -42.4346], [-0.3576, 1.0615, 1.5949, -37.1158],
                  [-0.2175, -1.8751, 6.9640, -26.6970]], np.float32)   # apps/camera_pipe/process.cpp:43-49
M7000 = np.array([[2.2997, -0.4478, 0.1706, -39.0923], [-0.3826, 1.5906, -0.2080, -25.4311],
                  [-0.0888, -0.7344, 2.2832, -20.0826]], np.float32)


def raw_frame(h, w, seed, bits=10):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = 400 + 300 * np.sin(xx / 23.0) * np.cos(yy / 17.0)
    noise = rng.integers(0, 1 << bits, (h, w))
    return np.clip(0.5 * base + 0.5 * noise, 0, (1 << bits) - 1).astype(np.uint16)


def run(hb, raw, out_shape, color_temp=3700.0, gamma=2.0, contrast=50.0, sharpen=1.0, black=25, white=1023, in_mins=None,
        out_mins=None):
    out = np.zeros(out_shape, np.uint8)
    bi = hb.HalideBuffer.from_numpy(raw, in_mins)
    b32, b70 = hb.HalideBuffer.from_numpy(M3200.copy()), hb.HalideBuffer.from_numpy(M7000.copy())
    bo = hb.HalideBuffer.from_numpy(out, out_mins, host_dirty=False)
    hb.filters.camera_pipe(bi, b32, b70, color_temp, gamma, contrast, sharpen, black, white, bo)
    bo.copy_to_host()
    return out


def harness_out_shape(h, w):
    return (3, ((h - 24) // 32) * 32, ((w - 32) // 32) * 32)  # apps/camera_pipe/process.cpp:34


@pytest.mark.parametrize("h,w", [(56, 64), (88, 96), (120, 200), (152, 288)])
@pytest.mark.parametrize("seed", [0, 1])
def test_matches_oracle_harness_shapes(hb, oracle, h, w, seed):
    raw = raw_frame(h, w, seed)
    shape = harness_out_shape(h, w)
    got = run(hb, raw, shape)
    want = oracle.camera_pipe(raw, M3200, M7000, 3700.0, 2.0, 50.0, 1.0, 25, 1023, shape)
    assert np.array_equal(got, want)


def test_full_range_raw_and_wraparound(hb, oracle):
    """16-bit noise: the u16 correction terms wrap and the i16 reinterpretation goes negative."""
    rng = np.random.default_rng(5)
    raw = rng.integers(0, 65536, (120, 160), dtype=np.uint16)
    shape = harness_out_shape(120, 160)
    assert np.array_equal(run(hb, raw, shape), oracle.camera_pipe(raw, M3200, M7000, 3700.0, 2.0, 50.0, 1.0, 25, 1023, shape))


@pytest.mark.parametrize("color_temp,gamma,contrast,sharpen,black,white",
                         [(3200.0, 1.0, 0.0, 0.0, 0, 1023), (7000.0, 2.2, 100.0, 3.9, 64, 900), (5000.0, 1.8, 25.0, 8.5, 25, 1023)])
def test_parameter_sweep(hb, oracle, color_temp, gamma, contrast, sharpen, black, white):
    raw = raw_frame(88, 128, 3)
    shape = harness_out_shape(88, 128)
    got = run(hb, raw, shape, color_temp, gamma, contrast, sharpen, black, white)
    want = oracle.camera_pipe(raw, M3200, M7000, color_temp, gamma, contrast, sharpen, black, white, shape)
    assert np.array_equal(got, want)


def test_odd_output_offsets(hb, oracle):
    """Output origin at odd coordinates: Bayer parity follows the absolute coordinates."""
    raw = raw_frame(140, 180, 9)
    kw = dict(out_mins=(3, 5, 0))
    got = run(hb, raw, (3, 70, 90), **kw)
    want = oracle.camera_pipe(raw, M3200, M7000, 3700.0, 2.0, 50.0, 1.0, 25, 1023, (3, 70, 90), **kw)
    assert np.array_equal(got, want)


def test_input_too_small_is_rejected(hb):
    from halide_b200 import HalideError
    raw = raw_frame(64, 64, 0)
    with pytest.raises(HalideError) as e:
        run(hb, raw, (3, 64, 64))
    assert e.value.code == -4


def test_harness_frame_size(hb, oracle):
    """2592x1968 raw -> 2560x1920x3 (the reference harness size); the oracle needs ~1 s for it."""
    raw = raw_frame(1968, 2592, 11)
    shape = harness_out_shape(1968, 2592)
    assert shape == (3, 1920, 2560)
    got = run(hb, raw, shape)
    want = oracle.camera_pipe(raw, M3200, M7000, 3700.0, 2.0, 50.0, 1.0, 25, 1023, shape)
    assert np.array_equal(got, want)
