"""GPU tests of the runtime's copies (reference: src/runtime/device_buffer_utils.h, cuda.cpp:884-1017,
device_interface.cpp:154-205): only the elements a buffer describes move — never the gaps between the rows or planes of
a padded / cropped buffer — and halide_buffer_copy copies a region between buffers of different layouts."""
import ctypes

import numpy as np
import pytest

from util import u16_frame

pytestmark = pytest.mark.gpu


def _dirty_pool(hb, nbytes):
    """Leave a freed 0xFF-filled block of exactly `nbytes` in the runtime's allocation pool, so that the next device
    allocation of that size starts out with non-zero garbage (a fresh cudaMalloc is often zero)."""
    from halide_b200 import HalideBuffer
    junk = np.full(nbytes, 0xFF, np.uint8)
    b = HalideBuffer.from_numpy(junk)
    hb.lib.check(hb.capi.halide_copy_to_device(None, b.ptr, ctypes.c_void_p(hb.capi.halide_cuda_device_interface())))
    b.device_free()


def test_copy_to_host_of_a_padded_output_leaves_the_padding_alone(hb, oracle):
    """Row stride > extent: the columns past the buffer's extent belong to the caller.  The device allocation of such a
    buffer spans them too; copy_to_host must not bring its (uninitialised) contents back over them."""
    from halide_b200 import HalideBuffer, filters
    inp = u16_frame((50, 70), 7)
    out_store = np.full((40, 64), 0xABCD, np.uint16)   # 61 columns used, 3 of padding per row
    out = out_store[:, :61]
    _dirty_pool(hb, ((40 - 1) * 64 + 61) * 2)
    bi = HalideBuffer.from_numpy(inp, mins=(-3, 10))
    bo = HalideBuffer.from_numpy(out, mins=(4, 15), host_dirty=False)
    filters.halide_blur(bi, bo)
    bo.copy_to_host()
    want = oracle.blur(inp, out_shape=(40, 61), in_mins=(-3, 10), out_mins=(4, 15))
    assert np.array_equal(out, want)
    assert (out_store[:, 61:] == 0xABCD).all()   # padding columns untouched


def test_copy_to_host_of_a_planar_crop_leaves_the_rest_of_the_image_alone(hb, oracle):
    """Output = a crop (rows, columns and channels) of a larger host image: everything outside the crop keeps its
    host values."""
    from halide_b200 import HalideBuffer, filters
    img = u16_frame((3, 90, 120), 4)
    canvas = np.full((3, 90, 120), 0x1234, np.uint16)
    crop = canvas[:, 21:71, 10:70]        # 50 rows x 60 columns at (x, y) = (10, 21)
    _dirty_pool(hb, crop.size * 0 + ((3 - 1) * 90 * 120 + (50 - 1) * 120 + 60) * 2)
    bi = HalideBuffer.from_numpy(img)
    bo = HalideBuffer.from_numpy(crop, mins=(10, 21, 0), host_dirty=False)
    filters.local_laplacian(bi, 8, 1.0 / 7.0, 1.0, bo)
    bo.copy_to_host()
    want = oracle.local_laplacian(img, 8, 1.0 / 7.0, 1.0, out_shape=(3, 50, 60), in_mins=(0, 0, 0), out_mins=(10, 21, 0))
    assert np.array_equal(crop, want)
    mask = np.ones_like(canvas, bool)
    mask[:, 21:71, 10:70] = False
    assert (canvas[mask] == 0x1234).all()


def test_strided_input_upload(hb, oracle):
    """Input = a padded host image (row stride > extent): the upload is strided as well, and the result is unaffected
    by what the padding holds."""
    from halide_b200 import HalideBuffer, filters
    store = u16_frame((52, 80), 9)
    inp = store[:, :72]
    out = np.zeros((50, 70), np.uint16)
    bi, bo = HalideBuffer.from_numpy(inp), HalideBuffer.from_numpy(out, host_dirty=False)
    filters.halide_blur(bi, bo)
    bo.copy_to_host()
    assert np.array_equal(out, oracle.blur(np.ascontiguousarray(inp)))


def test_buffer_copy_between_layouts(hb):
    """halide_buffer_copy: region of a device-resident source -> host buffer of another layout, host -> device, and
    device -> device; only dst's region is written, dirty bits follow device_interface.cpp:154-205."""
    from halide_b200 import HalideBuffer
    iface = ctypes.c_void_p(hb.capi.halide_cuda_device_interface())
    src_np = u16_frame((3, 40, 64), 2)
    src = HalideBuffer.from_numpy(src_np, mins=(5, -2, 0))
    hb.lib.check(hb.capi.halide_copy_to_device(None, src.ptr, iface))
    src_np_copy = src_np.copy()
    src_np[:] = 0          # the device copy is now the only valid one ...
    src.set_host_dirty(False)
    src.buf.flags |= 2     # ... and says so (device_dirty)
    # device -> host, into the middle of a padded canvas
    canvas = np.full((3, 30, 50), 7, np.uint16)
    dst_view = canvas[:, 4:24, 8:40]                      # 20 rows x 32 columns
    dst = HalideBuffer.from_numpy(dst_view, mins=(11, 3, 0), host_dirty=False)
    hb.lib.check(hb.capi.halide_buffer_copy(None, src.ptr, None, dst.ptr))
    assert np.array_equal(dst_view, src_np_copy[:, 5:25, 6:38])   # x: 11-5 = 6, y: 3-(-2) = 5
    mask = np.ones_like(canvas, bool)
    mask[:, 4:24, 8:40] = False
    assert (canvas[mask] == 7).all() and dst.host_dirty and not dst.device_dirty
    # device -> device (another allocation), then back to the host through copy_to_host
    out_np = np.zeros((2, 10, 16), np.uint16)
    d2 = HalideBuffer.from_numpy(out_np, mins=(20, 10, 1), host_dirty=False)
    hb.lib.check(hb.capi.halide_buffer_copy(None, src.ptr, iface, d2.ptr))
    assert d2.device_dirty
    d2.copy_to_host()
    assert np.array_equal(out_np, src_np_copy[1:3, 12:22, 15:31])
    # a region outside the source is an error (-4), not a silent clamp
    far = HalideBuffer.from_numpy(np.zeros((3, 10, 16), np.uint16), mins=(60, 0, 0), host_dirty=False)
    assert hb.capi.halide_buffer_copy(None, src.ptr, None, far.ptr) == -4


def test_copy_error_codes(hb):
    """copy_to_host with both dirty bits set -> -37 (host_and_device_dirty); device-dirty buffer without a host
    pointer -> -34 (host_is_null) (test/generator/error_codes_aottest.cpp, src/runtime/device_interface.cpp:30-56)."""
    from halide_b200 import HalideBuffer
    iface = ctypes.c_void_p(hb.capi.halide_cuda_device_interface())
    a = np.zeros((8, 8), np.uint16)
    b = HalideBuffer.from_numpy(a)
    hb.lib.check(hb.capi.halide_copy_to_device(None, b.ptr, iface))
    b.buf.flags = 3
    assert hb.capi.halide_copy_to_host(None, b.ptr) == -37
    b.buf.flags = 2
    host = b.buf.host
    b.buf.host = None
    assert hb.capi.halide_copy_to_host(None, b.ptr) == -34
    b.buf.host = host
    b.buf.flags = 0
