"""Multi-GPU parity worker (launched by torchrun, one rank per GPU): every rank runs the row-sharded
local_laplacian on its band and compares it bit-exactly with the single-GPU filter run on the whole frame on
its own GPU.   torchrun --nproc-per-node N tools/dist_check.py [W] [band_h]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as td  # noqa: E402


def main():
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    band_h = int(sys.argv[2]) if len(sys.argv) > 2 else 640
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = td.get_rank(), td.get_world_size()
    import halide_b200
    from halide_b200 import HalideBuffer, dist, filters
    halide_b200.capi.halide_b200_set_device(local)
    H = band_h * world
    rng = np.random.default_rng(123)
    frame = rng.integers(0, 65536, (3, H, W), dtype=np.uint16)
    # single-GPU result on the whole frame
    full_out = np.zeros_like(frame)
    bi, bo = HalideBuffer.from_numpy(frame), HalideBuffer.from_numpy(full_out, host_dirty=False)
    filters.local_laplacian(bi, 8, 1.0 / 7.0, 1.0, bo)
    bo.copy_to_host()
    # sharded result on this rank's band
    sh = dist.RowSharder(rank, world, W, band_h)
    band_in = np.ascontiguousarray(frame[:, sh.lo:sh.hi + 1, :])
    band_out = np.zeros_like(band_in)
    b_in, b_out = HalideBuffer.from_numpy(band_in), HalideBuffer.from_numpy(band_out, host_dirty=False)
    want = full_out[:, sh.lo:sh.hi + 1, :]
    bad = 0
    # gathered level: chosen by size (0), forced levels
    modes = [int(m) for m in os.environ.get("DIST_CHECK_MODES", "0,2,3,6").split(",")]
    for mode in modes:
        halide_b200.capi.halide_b200_ll_shard_coarse_level(mode)
        for it in range(3):  # repeated: pooled scratch, epochs, ready/gather handshakes of consecutive calls
            band_out[:] = 0
            b_in.set_host_dirty(True)
            sh.local_laplacian(b_in, 8, 1.0 / 7.0, 1.0, b_out)
            b_out.copy_to_host()
            n = int((band_out != want).sum())
            if n:
                print(f"rank {rank}: mode {mode} call {it}: {n} mismatching samples", flush=True)
            bad += n
    halide_b200.capi.halide_b200_ll_shard_coarse_level(0)
    t = torch.tensor([bad], device="cuda")
    td.all_reduce(t)
    if rank == 0:
        print(f"DIST_CHECK world={world} frame={W}x{H} band_h={band_h} mismatches={int(t.item())}")
    td.barrier()
    td.destroy_process_group()
    sys.exit(0 if int(t.item()) == 0 else 1)


if __name__ == "__main__":
    main()
