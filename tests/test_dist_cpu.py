"""CPU tests of the multi-GPU host logic: band geometry (which pyramid rows each rank owns / holds)
and a world_size-2 gloo run of the control-plane code in halide_b200/dist.py."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("frame_h,world", [(4320, 2), (2160 * 8, 8), (16384, 8), (2048, 2), (3000, 3), (16384, 4), (16384, 5), (8191, 7),
                                           (32768, 16)])
def test_shard_rows_cover_every_tap(hb, frame_h, world):
    """Exchange-free row sharding (ll_geom.h: ShardLevel): the `own` rows partition every level, and the rows a rank
    computes (d) cover every tap of the rows it needs one level up, so nothing but the input halo and the gathered level
    ever crosses a shard boundary."""
    from halide_b200 import dist
    w = 3840
    jr = dist.shard_plan_level(w, frame_h, world)
    assert 2 <= jr <= 7
    per_rank, needs = [], []
    for r in range(world):
        lo, hi = dist.band_rows(r, world, frame_h)
        lv, need = dist.band_geometry(w, frame_h, lo, hi, r == 0, r == world - 1, jr)
        per_rank.append(lv)
        needs.append((lo, hi, need))
    for j in range(1, 8):
        S = per_rank[0][j]["S"]
        # owned rows partition the whole frame's rows of that level, in rank order, without gaps
        assert per_rank[0][j]["own"][0] == S[0] and per_rank[-1][j]["own"][1] == S[1]
        for r in range(world - 1):
            assert per_rank[r][j]["own"][1] + 1 == per_rank[r + 1][j]["own"][0], (j, r)
    for r in range(world):
        lo, hi, need = needs[r]
        lv = per_rank[r]
        assert lv[0]["u"] == (lo, hi)
        for j in range(1, 8):
            S, d, u = lv[j]["S"], lv[j]["d"], lv[j]["u"]
            if j >= jr:
                assert d == S   # replicated levels are held whole
                continue
            # outGPyramid[j] rows needed = the bilinear taps (y-1)//2, (y+1)//2 of the rows needed one level down
            below = lv[j - 1]["u"]
            assert u == ((below[0] - 1) // 2, (below[1] + 1) // 2)
            # the rows computed cover the needed rows (inside the level) ...
            assert d[0] <= max(u[0], S[0]) and min(u[1], S[1]) <= d[1]
            # ... and the 1-3-3-1 taps 2y-1 .. 2y+2 of every row this rank computes one level up
            nxt = lv[j + 1]["own"] if j + 1 == jr else lv[j + 1]["d"]
            assert d[0] <= max(2 * nxt[0] - 1, S[0]) and min(2 * nxt[1] + 2, S[1]) <= d[1], (r, j)
        # input rows read: the taps of level 1, clipped to the frame; the halo beyond the band is small and comes
        # from the direct neighbours only
        c1 = lv[1]["own"] if jr == 1 else lv[1]["d"]
        assert need[0] <= max(2 * c1[0] - 1, 0) and min(2 * c1[1] + 2, frame_h - 1) <= need[1]
        assert need[0] >= (0 if r == 0 else needs[r - 1][0]) and need[1] <= (frame_h - 1 if r == world - 1 else needs[r + 1][1])
        assert lo - need[0] <= 2 ** (jr + 1) and need[1] - hi <= 2 ** (jr + 2)


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
        geo, need = dist.band_geometry(3840, 4320, lo, hi, r == 0, r == w - 1, dist.shard_plan_level(3840, 4320, w))
        allg = [None] * w
        td.all_gather_object(allg, (lo, hi, geo[1]["own"][0], geo[1]["own"][1]))
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
    """The gathered level depends only on the frame and the rank count (every rank must pick the same one): the level
    that minimises the bytes a rank receives per call — input halo (~3 * 2^j frame rows) plus (N-1)/N of the level; the
    override forces a level."""
    lvl = hb.capi.halide_b200_ll_shard_plan_level
    assert lvl(16384, 16384, 8) == 5 and lvl(16384, 16384, 4) == 5 and lvl(16384, 16384, 2) == 5
    assert lvl(3840, 2160 * 2, 2) == 4
    assert lvl(3840, 2160 * 8, 8) == 5
    assert lvl(1000, 1280, 2) == 4

    def received(w, h, n, j):   # the cost model itself, restated: halo + share of the level, in bytes
        halo = 3 * (1 << j) * w * 6
        level = (h // (1 << j) + 3) * (w // (1 << j) + 3) * 9 * 4
        return halo + level * (n - 1) / n
    for (w, h, n) in [(16384, 16384, 8), (3840, 4320, 2), (7680, 4320, 4)]:
        j = lvl(w, h, n)
        assert 2 <= j <= 7
        assert received(w, h, n, j) <= 1.1 * min(received(w, h, n, k) for k in range(2, 8))
    hb.capi.halide_b200_ll_shard_coarse_level(6)
    assert lvl(3840, 2160 * 2, 2) == 6
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
