"""Host logic of halide_b200.FramePipeline (no GPU: device=None skips the per-worker CUDA stream setup)."""
import threading
import time

import pytest


def test_slots_run_in_order_on_one_worker_and_results_round_trip(hb):
    seen = {0: [], 1: [], 2: []}
    workers = {0: set(), 1: set(), 2: set()}

    def job(slot, k):
        workers[slot].add(threading.get_ident())
        time.sleep(0.001 * ((k * 7) % 3))
        seen[slot].append(k)
        return slot * 1000 + k

    with hb.FramePipeline(depth=3, device=None) as fp:
        tickets = [(fp.submit(job, k % 3, k, slot=k % 3), k) for k in range(30)]
        for t, k in reversed(tickets):  # results can be collected in any order
            assert fp.result(t) == (k % 3) * 1000 + k
    for slot in range(3):
        assert seen[slot] == sorted(seen[slot]) and len(seen[slot]) == 10  # submission order within a slot
        assert len(workers[slot]) == 1                                     # one worker (hence one stream) per slot
    assert len(set.union(*workers.values())) == 3


def test_exceptions_surface_at_result_and_do_not_kill_the_worker(hb):
    def boom():
        raise ValueError("bad frame")

    fp = hb.FramePipeline(depth=1, device=None)
    t_bad, t_ok = fp.submit(boom), fp.submit(lambda: 42)
    with pytest.raises(ValueError, match="bad frame"):
        fp.result(t_bad)
    assert fp.result(t_ok) == 42
    fp.close()
    fp.close()  # idempotent


def test_depth_must_be_positive(hb):
    with pytest.raises(ValueError):
        hb.FramePipeline(depth=0, device=None)
