"""Drop-in proof: the reference's UNMODIFIED harness sources (apps/blur/test.cpp,
apps/*/process.cpp, apps/bilateral_grid/filter.cpp) compiled against include/*.h and linked against
libhalide_b200.so (built by oracle/Makefile into oracle/_ref/ while /root/reference is present; the
binaries travel to the GPU box, the sources do not).  Each must print "Success!" (what the
reference's ctest greps, apps/*/CMakeLists.txt) and its saved output must equal the oracle."""
import os
import subprocess

import numpy as np
import pytest

from util import f32_frame, u16_frame

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _bin(name):
    p = os.path.join(REF, name)
    if not os.path.exists(p):
        pytest.skip(f"{p} not built (needs /root/reference at build time)")
    return p


def _save_halide_npy(path, arr):
    """tools/halide_image_io.h:1433-1445 writes the Halide extents (x, y, c) as the numpy shape over
    x-fastest data; build the same file from an array indexed [c, y, x]."""
    shape = tuple(reversed(arr.shape))
    np.save(path, np.ascontiguousarray(arr).reshape(-1).reshape(shape))


def _load_halide_npy(path):
    """Reads a .npy written by tools/halide_image_io.h:1433-1475.  The reference pads the header AFTER its
    trailing newline, which numpy's own loader rejects, so the v1 header is parsed by hand.  The shape holds
    the Halide extents (x, y, c); the payload is x-fastest."""
    import ast
    import struct
    raw = open(path, "rb").read()
    assert raw[:6] == b"\x93NUMPY"
    (hlen,) = struct.unpack("<H", raw[8:10])
    meta = ast.literal_eval(raw[10:10 + hlen].decode("latin1").strip())
    assert not meta["fortran_order"]
    data = np.frombuffer(raw[10 + hlen:], dtype=np.dtype(meta["descr"]))
    return data.reshape(tuple(reversed(meta["shape"])))


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Success!" in r.stdout, r.stdout + r.stderr
    return r.stdout


def test_blur_test_cpp_passes_unmodified():
    """apps/blur/test.cpp:157-195 compares our halide_blur with the reference's two C implementations
    on a 2568x1922 12-bit frame and aborts on any difference."""
    out = _run([_bin("blur_test")])
    assert "times:" in out


def test_local_laplacian_process_cpp(tmp_path, oracle):
    img = u16_frame((3, 120, 200), 12)
    _save_halide_npy(tmp_path / "in.npy", img)
    out = _run([_bin("ll_process"), str(tmp_path / "in.npy"), "8", "1", "1", "3", str(tmp_path / "out.npy")])
    assert "Manually-tuned time" in out and "Auto-scheduled time" in out
    got = _load_halide_npy(tmp_path / "out.npy")
    # process.cpp:31 passes alpha / (levels - 1) computed in float
    want = oracle.local_laplacian(img, 8, float(np.float32(1.0) / np.float32(7)), 1.0)
    assert np.array_equal(got, want)


def test_stencil_chain_process_cpp(tmp_path, oracle):
    img = u16_frame((3, 70, 90), 13)
    _save_halide_npy(tmp_path / "in.npy", img)
    _run([_bin("sc_process"), str(tmp_path / "in.npy"), "3", str(tmp_path / "out.npy")])
    got = _load_halide_npy(tmp_path / "out.npy")
    assert np.array_equal(got, oracle.stencil_chain(np.ascontiguousarray(img[0])))  # harness takes the red channel


def test_bilateral_grid_filter_cpp(tmp_path, oracle):
    img = f32_frame((96, 136), 14)
    _save_halide_npy(tmp_path / "in.npy", img)
    _run([_bin("bg_filter"), str(tmp_path / "in.npy"), str(tmp_path / "out.npy"), "0.1", "3"])
    got = _load_halide_npy(tmp_path / "out.npy")
    want = oracle.bilateral_grid(img, 0.1)
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-6)) <= 1e-4


def test_nl_means_process_cpp(tmp_path, oracle):
    img = f32_frame((3, 40, 56), 15)
    _save_halide_npy(tmp_path / "in.npy", img)
    _run([_bin("nl_process"), str(tmp_path / "in.npy"), "7", "7", "0.12", "3", str(tmp_path / "out.npy")])
    got = _load_halide_npy(tmp_path / "out.npy")
    want = oracle.nl_means(img, 7, 7, 0.12)
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)) <= 1e-4


def test_camera_pipe_process_cpp(tmp_path, oracle):
    from test_camera_pipe_gpu import M3200, M7000, raw_frame
    raw = raw_frame(152, 224, 21)
    _save_halide_npy(tmp_path / "raw.npy", raw)
    _run([_bin("cp_process"), str(tmp_path / "raw.npy"), "3700", "2.0", "50", "1.0", "3", str(tmp_path / "out.npy")])
    got = _load_halide_npy(tmp_path / "out.npy")
    shape = (3, ((152 - 24) // 32) * 32, ((224 - 32) // 32) * 32)
    want = oracle.camera_pipe(raw, M3200, M7000, 3700.0, 2.0, 50.0, 1.0, 25, 1023, shape)
    assert np.array_equal(got, want)


def test_conv_layer_process_cpp():
    out = _run([_bin("conv_process")])
    assert "Manually-tuned time" in out and "Auto-scheduled time" in out


@pytest.mark.parametrize("name", ["local_laplacian", "bilateral_grid", "stencil_chain", "conv_layer", "nl_means", "camera_pipe"])
def test_rungen_benchmark_mode(name):
    """the reference's generic driver (tools/RunGenMain.cpp, unmodified) linked with our registration
    TU sizes its buffers through our bounds-query mode, fills them from the metadata estimates and times the filter —
    the command `ctest -L benchmark_apps` runs (apps/local_laplacian/CMakeLists.txt:62-70)."""
    r = subprocess.run([_bin(name + "_rungen"), "--benchmarks=all", "--estimate_all", "--parsable_output"], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "BEST_TIME_MSEC_PER_ITER" in r.stdout, r.stdout + r.stderr


def test_rungen_blur_explicit_extents(tmp_path):
    """halide_blur declares no estimates (apps/blur/halide_blur_generator.cpp has no set_estimates), so the generic driver
    is given the output extents and lets the bounds query size the input (`random:0:auto`, doc/RunGen.md:144-166); the
    result is written out and must equal the oracle on the same seeded input only up to RunGen's own RNG, so the check
    here is the driver's success + the timing line, and — reading the saved buffers back — blur's fixed point on the
    constant frame the second invocation feeds it."""
    exe = _bin("halide_blur_rungen")
    r = subprocess.run([exe, "--benchmarks=all", "--parsable_output", "--output_extents=[1920,1080]", "input=random:0:auto"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "BEST_TIME_MSEC_PER_ITER" in r.stdout, r.stdout + r.stderr
    out = tmp_path / "out.npy"
    r = subprocess.run([exe, "--output_extents=[64,48]", "input=constant:777:auto", f"blur_y={out}"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = _load_halide_npy(out)
    assert got.shape == (48, 64) and (got == 777).all()
