"""Row-sharded multi-GPU execution: one process per GPU (torchrun), NCCL halo exchange inside the library.

torch.distributed is only the control plane here (broadcast of the NCCL unique id, barriers in
bench.py); the data-path exchange — one ncclGroup of point-to-point halo messages per pyramid level
— is issued by libhalide_b200.so on its compute stream (halide_b200/csrc/hb_dist.cu).
"""
import ctypes
import os

from .lib import lib, check


def band_rows(rank, world, frame_h):
    """Rows [lo, hi] of the frame owned by `rank`: contiguous, top to bottom, balanced."""
    base, rem = divmod(frame_h, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0) - 1
    return lo, hi


def band_geometry(frame_w, frame_h, lo, hi, first, last):
    """Per-level rows owned / held by a band (host-only probe of ll_geom.h): list of 8 dicts."""
    out = (ctypes.c_int32 * 64)()
    check(lib.halide_b200_ll_band_geometry(frame_w, frame_h, lo, hi, int(first), int(last), out))
    keys = ("own_lo", "own_hi", "stored_lo", "stored_hi", "own_o_lo", "own_o_hi", "stored_o_lo", "stored_o_hi")
    return [dict(zip(keys, out[j * 8:(j + 1) * 8])) for j in range(8)]


_initialised = False


def init_from_torch_distributed():
    """Join the library's NCCL communicator using torch.distributed as the control plane."""
    global _initialised
    if _initialised:
        return
    import torch
    import torch.distributed as td
    rank, world = td.get_rank(), td.get_world_size()
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        check(lib.halide_b200_dist_unique_id(buf))
    if td.get_backend() == "nccl":
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
        td.broadcast(t, 0)
        raw = bytes(t.cpu().numpy().tobytes())
    else:
        obj = [buf.raw]
        td.broadcast_object_list(obj, 0)
        raw = obj[0]
    check(lib.halide_b200_dist_init(rank, world, raw))
    _initialised = True


class RowSharder:
    """Each rank holds a band of W x band_h rows of a frame of W x (band_h * world) rows."""

    def __init__(self, rank, world, width, band_h):
        self.rank, self.world, self.width, self.band_h = rank, world, width, band_h
        self.frame_h = band_h * world
        self.lo, self.hi = band_rows(rank, world, self.frame_h)
        init_from_torch_distributed()
        lvl = os.environ.get("HALIDE_B200_SHARD_COARSE_LEVEL")  # experiment knob; see halide_b200_ll_shard_coarse_level
        if lvl is not None:
            lib.halide_b200_ll_shard_coarse_level(int(lvl))

    def set_band_mins(self, buf):
        """Put a band buffer (rows 0..band_h-1 locally) at its rows of the frame."""
        buf.dims[1].min = self.lo

    def local_laplacian(self, in_band, levels, alpha, beta, out_band):
        self.set_band_mins(in_band)
        self.set_band_mins(out_band)
        return check(lib.halide_b200_local_laplacian_sharded(in_band.ptr, ctypes.c_int32(levels), ctypes.c_float(alpha),
                                                             ctypes.c_float(beta), out_band.ptr, ctypes.c_int32(0),
                                                             ctypes.c_int32(self.frame_h)))
