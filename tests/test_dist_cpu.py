"""CPU tests of the multi-GPU host logic: band geometry (which pyramid rows each rank owns / holds)
and a world_size-2 gloo run of the control-plane code in halide_b200/dist.py."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("frame_h,world", [(4320, 2), (2160 * 8, 8), (16384, 8), (2048, 2), (3000, 3)])
def test_bands_tile_every_level(hb, frame_h, world):
    from halide_b200 import dist
    w = 3840
    whole = dist.band_geometry(w, frame_h, 0, frame_h - 1, True, True)
    per_rank = []
    for r in range(world):
        lo, hi = dist.band_rows(r, world, frame_h)
        per_rank.append(dist.band_geometry(w, frame_h, lo, hi, r == 0, r == world - 1))
    for j in range(1, 8):
        # owned rows partition the whole frame's rows of that level, in rank order, without gaps
        assert per_rank[0][j]["own_lo"] == whole[j]["own_lo"]
        assert per_rank[-1][j]["own_hi"] == whole[j]["own_hi"]
        for r in range(world - 1):
            assert per_rank[r][j]["own_hi"] + 1 == per_rank[r + 1][j]["own_lo"], (j, r)
            assert per_rank[r][j]["own_o_hi"] + 1 == per_rank[r + 1][j]["own_o_lo"], (j, r)
            # halo: one row above, two below on the Gaussian side; one and one on the output side
            assert per_rank[r][j]["stored_hi"] == per_rank[r][j]["own_hi"] + 2
            assert per_rank[r + 1][j]["stored_lo"] == per_rank[r + 1][j]["own_lo"] - 1
            assert per_rank[r][j]["stored_o_hi"] == per_rank[r][j]["own_o_hi"] + 1
            assert per_rank[r + 1][j]["stored_o_lo"] == per_rank[r + 1][j]["own_o_lo"] - 1
        # the taps of level j+1 (rows 2y-1 .. 2y+2) of every owned row stay inside the held rows of level j
        if j < 7:
            for r in range(world):
                g, gn = per_rank[r][j], per_rank[r][j + 1]
                lo_need = max(2 * gn["own_lo"] - 1, whole[j]["own_lo"])
                hi_need = min(2 * gn["own_hi"] + 2, whole[j]["own_hi"])
                assert g["stored_lo"] <= lo_need and hi_need <= g["stored_hi"], (j, r)


def test_band_rows_balanced():
    from halide_b200 import dist
    rows = [dist.band_rows(r, 3, 10) for r in range(3)]
    assert rows == [(0, 3), (4, 6), (7, 9)]


def test_gloo_world2_control_plane(tmp_path):
    """Two CPU processes over gloo: each computes its band, gathers everyone's, and checks the tiling —
    the same code path bench.py takes under torchrun, minus the NCCL communicator."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import torch.distributed as td
        from halide_b200 import dist
        td.init_process_group("gloo")
        r, w = td.get_rank(), td.get_world_size()
        lo, hi = dist.band_rows(r, w, 4320)
        geo = dist.band_geometry(3840, 4320, lo, hi, r == 0, r == w - 1)
        allg = [None] * w
        td.all_gather_object(allg, (lo, hi, geo[1]["own_lo"], geo[1]["own_hi"]))
        assert allg[0][1] + 1 == allg[1][0]
        assert allg[0][3] + 1 == allg[1][2]
        td.barrier()
        if r == 0:
            print("OK", allg)
        td.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout


def test_coarse_level_choice(hb):
    """The gathered level depends only on the frame and the rank count (every rank must pick the same one):
    4K bands -> level 4 (1.2 MB per peer) at N=2, 4 and 8; smaller frames gather a finer level; more ranks
    than flag slots, or the -1 override, fall back to level-by-level exchange (8 = no gather)."""
    lvl = hb.capi.halide_b200_ll_shard_plan_level
    assert lvl(3840, 2160 * 2, 2) == 4
    assert lvl(3840, 2160 * 4, 4) == 4
    assert lvl(3840, 2160 * 8, 8) == 4
    assert lvl(1000, 1280, 2) == 3
    assert lvl(3840, 2160, 1) == 8
    assert lvl(3840, 2160 * 16, 16) == 8
    hb.capi.halide_b200_ll_shard_coarse_level(-1)
    assert lvl(3840, 2160 * 2, 2) == 8
    hb.capi.halide_b200_ll_shard_coarse_level(5)
    assert lvl(3840, 2160 * 2, 2) == 5
    hb.capi.halide_b200_ll_shard_coarse_level(0)


@pytest.mark.parametrize("world", [2, 3])
def test_input_halo_sharding_matches_whole_frame_oracle(world):
    """SURVEY §8e for blur, nl_means, stencil_chain, bilateral_grid and camera_pipe: the sharding is host logic only
    (halo rule, row exchange over torch.distributed, buffer placement by mins), so it is checked end to end on CPU —
    gloo, CPU tensors, the oracle standing in for the single-GPU filter — against the oracle on the whole frame."""
    port = 29518 + world
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "tests", "dist_rows_worker.py")],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert f"ROWS_CHECK world={world} failures=[]" in out.stdout
