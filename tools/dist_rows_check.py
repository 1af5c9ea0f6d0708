"""Multi-GPU parity worker for the input-halo filters (torchrun, one rank per GPU): every rank runs the row-sharded
blur / nl_means / stencil_chain / bilateral_grid / camera_pipe on its band (halide_b200.dist.InputHaloSharder: NCCL row
exchange + the ordinary single-GPU filter) and compares it with the same filter run on the whole frame on its own GPU (bit-exact for the integer pipelines,
1e-4 relative for the float ones).
    torchrun --nproc-per-node N tools/dist_rows_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as td  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = td.get_rank(), td.get_world_size()
    import halide_b200
    from halide_b200 import HalideBuffer, dist, filters
    halide_b200.capi.halide_b200_set_device(local)
    dev = torch.device("cuda", local)
    sh = dist.InputHaloSharder(rank, world)
    rng = np.random.default_rng(11)
    bad = {}

    def u16(shape, hi=65536):
        return torch.from_numpy(rng.integers(0, hi, shape, dtype=np.uint16).view(np.int16)).to(dev).view(torch.uint16)

    def f32(shape):
        return torch.from_numpy(rng.random(shape, dtype=np.float32)).to(dev)

    def whole(run, inp, out):
        bi, bo = HalideBuffer.from_torch(inp), HalideBuffer.from_torch(out)
        run(bi, bo)
        bo.device_sync()

    def differs(a, b):
        if a.dtype == torch.float32:  # the float pipelines' parity bar (1e-4 relative): tiling may reorder a few sums
            return int((~torch.isclose(a, b, rtol=1e-4, atol=1e-6)).sum().item())
        if a.dtype == torch.uint16:
            a, b = a.view(torch.int16), b.view(torch.int16)
        return int((a != b).sum().item())

    def band_of(t, rows, row_dim=-2):
        return t.narrow(row_dim % t.dim(), rows[0], rows[1] - rows[0] + 1).contiguous()

    W = 640
    # blur
    H = 96 * world + 5
    img, full = u16((H + 2, W + 2)), torch.zeros((H, W), dtype=torch.uint16, device=dev)
    whole(lambda bi, bo: filters.halide_blur(bi, bo), img, full)
    rows = dist.band_rows(rank, world, H)
    own = dist.default_in_own(rank, world, rows, (0, H + 1))
    out = torch.zeros((rows[1] - rows[0] + 1, W), dtype=torch.uint16, device=dev)
    sh.halide_blur(band_of(img, own), out, H)
    bad["blur"] = differs(out, band_of(full, rows))
    # nl_means
    H = 64 * world + 3
    img, full = f32((3, H, W)), torch.zeros((3, H, W), dtype=torch.float32, device=dev)
    whole(lambda bi, bo: filters.nl_means(bi, 3, 7, 0.12, bo), img, full)
    rows = dist.band_rows(rank, world, H)
    out = torch.zeros((3, rows[1] - rows[0] + 1, W), dtype=torch.float32, device=dev)
    sh.nl_means(band_of(img, rows), 3, 7, 0.12, out, H)
    bad["nl_means"] = differs(out, band_of(full, rows))
    # stencil_chain
    H = 80 * world
    img, full = u16((H, W)), torch.zeros((H, W), dtype=torch.uint16, device=dev)
    whole(lambda bi, bo: filters.stencil_chain(bi, bo), img, full)
    rows = dist.band_rows(rank, world, H)
    out = torch.zeros((rows[1] - rows[0] + 1, W), dtype=torch.uint16, device=dev)
    sh.stencil_chain(band_of(img, rows), out, H)
    bad["stencil_chain"] = differs(out, band_of(full, rows))
    # bilateral_grid
    H = 72 * world + 1
    img, full = f32((H, W)), torch.zeros((H, W), dtype=torch.float32, device=dev)
    whole(lambda bi, bo: filters.bilateral_grid(bi, 0.1, bo), img, full)
    rows = dist.band_rows(rank, world, H)
    out = torch.zeros((rows[1] - rows[0] + 1, W), dtype=torch.float32, device=dev)
    sh.bilateral_grid(band_of(img, rows), 0.1, out, H)
    bad["bilateral_grid"] = differs(out, band_of(full, rows))
    # camera_pipe
    out_h, out_w = 64 * world, 512
    raw = u16((out_h + 56, out_w + 64), hi=1024)
    m32 = torch.from_numpy((rng.random((3, 4), dtype=np.float32) * 2 - 0.5).astype(np.float32))
    m70 = torch.from_numpy((rng.random((3, 4), dtype=np.float32) * 2 - 0.5).astype(np.float32))
    args = (3700.0, 2.0, 50.0, 1.0, 25, 1023)
    full = torch.zeros((3, out_h, out_w), dtype=torch.uint8, device=dev)
    b32, b70 = HalideBuffer.from_torch(m32), HalideBuffer.from_torch(m70)
    bi, bo = HalideBuffer.from_torch(raw), HalideBuffer.from_torch(full)
    filters.camera_pipe(bi, b32, b70, *args, bo)
    bo.device_sync()
    rows = dist.band_rows(rank, world, out_h)
    own = dist.default_in_own(rank, world, rows, (0, raw.shape[0] - 1))
    out = torch.zeros((3, rows[1] - rows[0] + 1, out_w), dtype=torch.uint8, device=dev)
    sh.camera_pipe(band_of(raw, own), raw.shape[0], m32, m70, *args, out, out_h)
    bad["camera_pipe"] = differs(out, band_of(full, rows))

    torch.cuda.synchronize()
    t = torch.tensor([bad[k] for k in sorted(bad)], device=dev)
    td.all_reduce(t)
    total = int(t.sum().item())
    if rank == 0:
        print("DIST_ROWS_CHECK world=%d %s mismatches=%d" % (world, dict(zip(sorted(bad), t.tolist())), total))
    td.barrier()
    td.destroy_process_group()
    sys.exit(0 if total == 0 else 1)


if __name__ == "__main__":
    main()
