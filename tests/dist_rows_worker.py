"""gloo worker for tests/test_dist_cpu.py::test_input_halo_sharding_*: every rank runs the row-sharded host logic of
halide_b200.dist (halo rule, row exchange, buffer placement) on CPU tensors, with the CPU oracle standing in for the
single-GPU filter, and compares its band with the oracle run on the whole frame."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as td  # noqa: E402

from halide_b200 import dist  # noqa: E402
from oracle import pyoracle  # noqa: E402


def main():
    td.init_process_group("gloo")
    rank, world = td.get_rank(), td.get_world_size()
    pyoracle.set_threads(2)
    rng = np.random.default_rng(5)
    W = 96
    failures = []

    def check(name, in_frame, want, oracle_call, halo, row_dim_in=-2):
        """in_frame: numpy, rows along axis -2; want: whole-frame oracle output (rows along -2)."""
        out_h, in_h = want.shape[-2], in_frame.shape[-2]
        lo, hi = dist.band_rows(rank, world, out_h)
        a, b = dist.default_in_own(rank, world, (lo, hi), (0, in_h - 1))
        band = torch.from_numpy(np.ascontiguousarray(in_frame[..., a:b + 1, :]))
        out_shape = list(want.shape)
        out_shape[-2] = hi - lo + 1
        out_band = torch.zeros(out_shape, dtype=torch.from_numpy(want[..., :1, :1].copy()).dtype)

        def call(ext, ext_row0, out_t, out_row0):
            in_mins = [0] * ext.dim()
            out_mins = [0] * out_t.dim()
            in_mins[1], out_mins[1] = ext_row0, out_row0
            res = oracle_call(ext.numpy(), tuple(out_t.shape), tuple(in_mins), tuple(out_mins))
            out_t.copy_(torch.from_numpy(res))

        need = dist.run_input_halo_sharded(call, band, (a, b), (0, in_h - 1), out_band, (lo, hi), halo, rank, world)
        got = out_band.numpy()
        ref = want[..., lo:hi + 1, :]
        ok = np.array_equal(got, ref) if got.dtype.kind in "ui" else np.allclose(got, ref, rtol=1e-6, atol=1e-7)
        if not ok:
            failures.append((name, rank, need))

    # blur: input 2 rows taller / 2 columns wider than the output
    H = 61
    img = rng.integers(0, 65536, (H + 2, W + 2), dtype=np.uint16)
    check("blur", img, pyoracle.blur(img),
          lambda ext, oshape, imins, omins: pyoracle.blur(ext, out_shape=oshape, in_mins=imins, out_mins=omins),
          dist.halo_rows("halide_blur"))
    # nl_means (clamped at the frame edge), patch 3 / search 7 -> 4-row halo
    H = 57
    f3 = rng.random((3, H, W), dtype=np.float32)
    check("nl_means", f3, pyoracle.nl_means(f3, 3, 7, 0.12),
          lambda ext, oshape, imins, omins: pyoracle.nl_means(ext, 3, 7, 0.12, out_shape=oshape, in_mins=imins, out_mins=omins),
          dist.halo_rows("nl_means", patch_size=3, search_area=7))
    # stencil_chain with 32 stages -> 64-row halo: bands must be >= 64 rows for nearest-neighbour ownership
    H = 72 * world
    u = rng.integers(0, 65536, (H, W), dtype=np.uint16)
    check("stencil_chain", u, pyoracle.stencil_chain(u),
          lambda ext, oshape, imins, omins: pyoracle.stencil_chain(ext, out_shape=oshape, in_mins=imins, out_mins=omins),
          dist.halo_rows("stencil_chain"))
    # bilateral_grid: 32-row halo (grid cells yi-2 .. yi+3)
    H = 45 * world
    f = rng.random((H, W), dtype=np.float32)
    check("bilateral_grid", f, pyoracle.bilateral_grid(f, 0.1),
          lambda ext, oshape, imins, omins: pyoracle.bilateral_grid(ext, 0.1, out_shape=oshape, in_mins=imins, out_mins=omins),
          dist.halo_rows("bilateral_grid"))
    # camera_pipe: required raw rows asked from the filter's own bounds query
    from halide_b200 import HalideBuffer
    raw = rng.integers(0, 1024, (32 * world + 56, 160), dtype=np.uint16)
    m32 = (rng.random((3, 4), dtype=np.float32) * 2 - 0.5).astype(np.float32)
    m70 = (rng.random((3, 4), dtype=np.float32) * 2 - 0.5).astype(np.float32)
    args = (3700.0, 2.0, 50.0, 1.0, 25, 1023)
    out_h, out_w = 32 * world, 96
    want = pyoracle.camera_pipe(raw, m32, m70, *args, (3, out_h, out_w))
    lo, hi = dist.band_rows(rank, world, out_h)
    a, b = dist.default_in_own(rank, world, (lo, hi), (0, raw.shape[0] - 1))
    out_band = torch.zeros((3, hi - lo + 1, out_w), dtype=torch.uint8)
    ob = HalideBuffer.from_torch(out_band, mins=(0, lo, 0))
    need = dist.camera_pipe_need_rows(ob, HalideBuffer.from_numpy(m32), HalideBuffer.from_numpy(m70), args)
    ext = dist.exchange_rows(torch.from_numpy(np.ascontiguousarray(raw[a:b + 1])), (a, b), need, rank, world)
    got = pyoracle.camera_pipe(ext.numpy(), m32, m70, *args, (3, hi - lo + 1, out_w), in_mins=(0, need[0]), out_mins=(0, lo, 0))
    if not np.array_equal(got, want[:, lo:hi + 1, :]):
        failures.append(("camera_pipe", rank, need))

    allf = [None] * world
    td.all_gather_object(allf, failures)
    td.barrier()
    if rank == 0:
        flat = [f for fs in allf for f in fs]
        print("ROWS_CHECK world=%d failures=%s" % (world, flat))
    td.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
