"""Device self-test: the shared-reciprocal division and magic-number conversions used by the fast local_laplacian
kernels must equal div.rn.f32 / cvt bit for bit on billions of operands from the pipeline's ranges."""
import ctypes

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 2])
def test_arith_shortcuts_are_exact(hb, seed):
    l = hb.load_library()
    l.halide_b200_selftest_arith.restype = ctypes.c_longlong
    l.halide_b200_selftest_arith.argtypes = [ctypes.c_ulonglong, ctypes.c_ulonglong]
    assert l.halide_b200_selftest_arith(1 << 32, seed) == 0
