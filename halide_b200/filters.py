"""Python call sites of the C-ABI filters, one function per exported symbol.

Same names, argument order and error behaviour as the generated AOT functions the reference's
harnesses call (apps/*/process.cpp, apps/blur/test.cpp): a negative halide_error_code_t becomes a
:class:`HalideError`.  Outputs are left device-dirty, exactly like a Halide GPU filter; call
``out.copy_to_host()`` to read them on the host.
"""
import ctypes

from .lib import lib, check


def halide_blur(input, blur_y):
    """apps/blur/halide_blur_generator.cpp: 3x3 box filter, uint16."""
    return check(lib.halide_blur(input.ptr, blur_y.ptr))


def local_laplacian(input, levels, alpha, beta, output):
    """apps/local_laplacian/local_laplacian_generator.cpp.  NOTE: like the harness
    (process.cpp:31) callers pass alpha already divided by (levels - 1)."""
    return check(lib.local_laplacian(input.ptr, ctypes.c_int32(levels), ctypes.c_float(alpha),
                                     ctypes.c_float(beta), output.ptr))
