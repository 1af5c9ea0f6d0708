"""ctypes mirror of halide_buffer_t and helpers to describe numpy / torch memory with it.

Layout facts: include/halide_b200_runtime.h (reference src/runtime/HalideRuntime.h:1657-1737).
Halide dimension 0 is the innermost (x); numpy/torch shapes are written outermost-first, so an
array indexed ``a[c, y, x]`` is the Halide buffer ``(x, y, c)`` with strides ``(1, W, W*H)`` —
the planar layout Halide::Runtime::Buffer uses by default (HalideBuffer.h:441-460).
"""
import ctypes

import numpy as np

from .lib import lib, check

HOST_DIRTY = 1
DEVICE_DIRTY = 2

_TYPE_CODES = {"i": 0, "u": 1, "f": 2}


class halide_type_t(ctypes.Structure):
    _fields_ = [("code", ctypes.c_uint8), ("bits", ctypes.c_uint8), ("reserved", ctypes.c_uint16)]


class halide_dimension_t(ctypes.Structure):
    _fields_ = [("min", ctypes.c_int32), ("extent", ctypes.c_int32), ("stride", ctypes.c_int32),
                ("flags", ctypes.c_uint32)]


class halide_buffer_t(ctypes.Structure):
    _fields_ = [("device", ctypes.c_uint64),
                ("device_interface", ctypes.c_void_p),
                ("host", ctypes.c_void_p),
                ("flags", ctypes.c_uint64),
                ("type", halide_type_t),
                ("dimensions", ctypes.c_int32),
                ("dim", ctypes.POINTER(halide_dimension_t)),
                ("padding", ctypes.c_void_p)]


assert ctypes.sizeof(halide_buffer_t) == 56
assert ctypes.sizeof(halide_dimension_t) == 16
assert ctypes.sizeof(halide_type_t) == 4


class HalideBuffer:
    """Owns a halide_buffer_t + its dim[] and keeps the backing memory alive."""

    def __init__(self, dtype, shape_xyz, strides=None, mins=None):
        dtype = np.dtype(dtype)
        n = len(shape_xyz)
        self.dtype = dtype
        self.dims = (halide_dimension_t * max(n, 1))()
        stride = 1
        for d in range(n):
            self.dims[d].min = 0 if mins is None else int(mins[d])
            self.dims[d].extent = int(shape_xyz[d])
            self.dims[d].stride = int(stride if strides is None else strides[d])
            stride *= int(shape_xyz[d])
        self.buf = halide_buffer_t()
        self.buf.type = halide_type_t(_TYPE_CODES[dtype.kind], dtype.itemsize * 8, 0)
        self.buf.dimensions = n
        self.buf.dim = ctypes.cast(self.dims, ctypes.POINTER(halide_dimension_t))
        self._keep = None
        self._wrapped = False

    # -- constructors -------------------------------------------------------------------------
    @classmethod
    def from_numpy(cls, arr, mins=None, host_dirty=True):
        """Describe a numpy array (indexed outermost-first) as a host buffer."""
        shape = tuple(reversed(arr.shape))
        strides = tuple(s // arr.itemsize for s in reversed(arr.strides))
        b = cls(arr.dtype, shape, strides, mins)
        b.buf.host = arr.ctypes.data
        b._keep = arr
        if host_dirty:
            b.buf.flags = HOST_DIRTY
        return b

    @classmethod
    def from_torch(cls, t, mins=None):
        """Describe a torch tensor: CPU tensors become host buffers (pinned memory is used as
        is), CUDA tensors are wrapped as device-resident buffers (halide_cuda_wrap_device_ptr)."""
        import torch
        np_dtype = {torch.uint16: np.uint16, torch.int16: np.int16, torch.uint8: np.uint8,
                    torch.float32: np.float32, torch.int32: np.int32}[t.dtype]
        shape = tuple(reversed(t.shape))
        strides = tuple(reversed(t.stride()))
        b = cls(np_dtype, shape, strides, mins)
        b._keep = t
        if t.is_cuda:
            check(lib.halide_cuda_wrap_device_ptr(None, ctypes.byref(b.buf), ctypes.c_uint64(t.data_ptr())))
            b._wrapped = True
        else:
            b.buf.host = t.data_ptr()
            b.buf.flags = HOST_DIRTY
        return b

    @classmethod
    def bounds_query(cls, dtype, ndim):
        """A buffer with null host and device: asks the filter for the required region."""
        return cls(dtype, (0,) * ndim)

    # -- protocol -----------------------------------------------------------------------------
    @property
    def ptr(self):
        return ctypes.byref(self.buf)

    def set_host_dirty(self, v=True):
        if v:
            self.buf.flags |= HOST_DIRTY
        else:
            self.buf.flags &= ~HOST_DIRTY

    @property
    def device_dirty(self):
        return bool(self.buf.flags & DEVICE_DIRTY)

    @property
    def host_dirty(self):
        return bool(self.buf.flags & HOST_DIRTY)

    def copy_to_host(self):
        check(lib.halide_copy_to_host(None, self.ptr))

    def device_sync(self):
        check(lib.halide_device_sync(None, self.ptr))

    def device_free(self):
        if self._wrapped:
            check(lib.halide_cuda_detach_device_ptr(None, self.ptr))
            self._wrapped = False
            self.buf.flags &= ~DEVICE_DIRTY
        elif self.buf.device:
            check(lib.halide_device_free(None, self.ptr))

    def shape(self):
        return [(self.dims[d].min, self.dims[d].extent, self.dims[d].stride) for d in range(self.buf.dimensions)]

    def __del__(self):
        try:
            self.device_free()
        except Exception:
            pass
