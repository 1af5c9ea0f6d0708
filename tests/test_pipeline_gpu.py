"""Several caller threads, one CUDA stream each (halide_b200.FramePipeline): every frame's result must be
the bits the oracle / a plain serial call produces, whatever the interleaving of copies, kernels and the
per-stream device pool."""
import numpy as np
import pytest

from util import run_blur, run_local_laplacian, u16_frame

pytestmark = pytest.mark.gpu


def test_pipelined_local_laplacian_matches_oracle(hb, oracle):
    shapes = [(3, 200, 333), (3, 129, 257), (3, 64, 512)]  # different sizes: the pool hands out different blocks
    frames = [u16_frame(shapes[i % 3], 100 + i) for i in range(12)]
    want = [oracle.local_laplacian(f, 8, 1.0 / 7, 1.0) for f in frames[:3]]
    serial = [run_local_laplacian(hb, f, 8, 1.0 / 7, 1.0) for f in frames]
    for i in range(3):
        assert np.array_equal(serial[i], want[i])
    with hb.FramePipeline(depth=3) as fp:
        for rounds in range(3):  # repeated so blocks recycle through the per-stream free lists
            tickets = [fp.submit(run_local_laplacian, hb, f, 8, 1.0 / 7, 1.0) for f in frames]
            got = [fp.result(t) for t in tickets]
            for g, s in zip(got, serial):
                assert np.array_equal(g, s)


def test_pipelined_mixed_filters_and_errors(hb, oracle):
    frames = [u16_frame((130 + 8 * i, 515), 7 + i) for i in range(6)]
    with hb.FramePipeline(depth=2) as fp:
        tickets = [fp.submit(run_blur, hb, f, (f.shape[0] - 2, f.shape[1] - 2)) for f in frames]
        bad = fp.submit(run_blur, hb, frames[0], (frames[0].shape[0], frames[0].shape[1]))  # needs rows the input lacks
        for t, f in zip(tickets, frames):
            assert np.array_equal(fp.result(t), oracle.blur(f))
        with pytest.raises(hb.HalideError):
            fp.result(bad)


def test_stream_create_destroy(hb):
    s = hb.capi.halide_b200_stream_create()
    assert s
    hb.capi.halide_b200_set_stream(s)
    assert hb.capi.halide_b200_get_stream() == s
    f = u16_frame((66, 264), 3)
    got = run_blur(hb, f, (64, 262))
    assert hb.capi.halide_b200_stream_destroy(s) == 0
    assert hb.capi.halide_b200_get_stream() is None
    assert got.any()
