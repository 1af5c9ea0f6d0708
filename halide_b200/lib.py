"""ctypes loader for libhalide_b200.so (the C-ABI drop-in library).

Fails loudly when the library is missing: there is no CPU or PyTorch fallback for the filters.
"""
import ctypes
import os
import threading
from contextlib import contextmanager

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhalide_b200.so")

_lib = None
_lock = threading.Lock()


class HalideError(RuntimeError):
    """A filter returned a negative halide_error_code_t (include/halide_b200_runtime.h)."""

    def __init__(self, code, message):
        super().__init__(f"halide error {code}: {message}")
        self.code = code
        self.message = message


_ERROR_HANDLER_T = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_char_p)
_last_error = threading.local()


def _record_error(_uc, msg):
    _last_error.msg = msg.decode("utf-8", "replace") if msg else ""


# Keep a reference so the callback is never garbage collected.
_handler_ref = _ERROR_HANDLER_T(_record_error)


def load_library():
    """Load the shared library once; raise ImportError with build instructions if absent."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C halide_b200/csrc`). halide_b200 has no CPU fallback.")
        l = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        l.halide_set_error_handler.restype = ctypes.c_void_p
        l.halide_set_error_handler.argtypes = [_ERROR_HANDLER_T]
        l.halide_b200_kernel_launch_count.restype = ctypes.c_uint64
        l.halide_b200_target.restype = ctypes.c_char_p
        l.halide_b200_last_kernel_ms.restype = ctypes.c_float
        l.halide_b200_set_stream.argtypes = [ctypes.c_void_p]
        l.halide_b200_get_stream.restype = ctypes.c_void_p
        l.halide_b200_stream_create.restype = ctypes.c_void_p
        l.halide_b200_stream_destroy.argtypes = [ctypes.c_void_p]
        l.halide_cuda_device_interface.restype = ctypes.c_void_p
        l.halide_b200_profile_report.argtypes = [ctypes.c_char_p, ctypes.c_int]
        # The reference's default handler aborts the process (posix_error_handler.cpp); from
        # Python we record the message and raise HalideError from the returned code instead.
        l.halide_set_error_handler(_handler_ref)
        _lib = l
        return _lib


class _LazyLib:
    def __getattr__(self, name):
        return getattr(load_library(), name)


lib = _LazyLib()


def last_error_message():
    return getattr(_last_error, "msg", "")


def check(code):
    """Raise HalideError for a negative return code."""
    if code != 0:
        raise HalideError(code, last_error_message())
    return code


@contextmanager
def capture_errors():
    """Context in which the last error message is reset (for tests that expect failures)."""
    _last_error.msg = ""
    yield _last_error


def profile(enable=True):
    lib.halide_b200_profile_enable(1 if enable else 0)


def profile_reset():
    lib.halide_b200_profile_reset()


def profile_report():
    """Return {kernel_name: (count, total_ms)} accumulated since the last reset."""
    n = lib.halide_b200_profile_report(None, 0)
    buf = ctypes.create_string_buffer(n + 16)
    lib.halide_b200_profile_report(buf, n + 16)
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.rsplit(" ", 2)
        out[name] = (int(cnt), float(ms))
    return out
