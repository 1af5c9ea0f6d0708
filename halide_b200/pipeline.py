"""Frame pipelining over the C ABI: several caller threads, one CUDA stream each.

A filter call with host buffers is H2D copy -> kernels -> (copy_to_host) D2H copy, all on the calling
thread's stream (``halide_b200_set_stream`` is thread-local).  One thread therefore leaves the PCIe link
idle while the kernels run and uses only one direction of it at a time.  A video-style caller that has
several frames in hand gets the link's full-duplex bandwidth by running ``depth`` frames concurrently,
each on its own thread and stream: frame i's D2H overlaps frame i+1's kernels and frame i+2's H2D.

This is host-side plumbing only (``threading`` + the thread-local stream of the C ABI; ctypes releases
the GIL for the duration of every library call).  The reference has the same facility in the form of
``halide_set_cuda_get_stream`` / per-thread user contexts (src/runtime/HalideRuntimeCuda.h:66-81).
"""
import queue
import threading

from .lib import lib, check


class FramePipeline:
    """Runs ``fn(*args)`` calls on ``depth`` worker threads, each bound to its own CUDA stream.

    ``submit`` returns a ticket; ``result(ticket)`` blocks until that call has finished (including whatever
    ``fn`` did to bring the result to the host) and re-raises its exception, if any.  Jobs submitted with the
    same ``slot`` run on the same worker, in order — use one slot per set of buffers.  ``device=None`` skips the
    CUDA setup (workers run ``fn`` as is): for callers that bind streams themselves, and for the host-logic tests.
    """

    def __init__(self, depth=3, device=0):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.depth = depth
        self._queues = [queue.Queue() for _ in range(depth)]
        self._done = {}
        self._cv = threading.Condition()
        self._next = 0
        self._threads = [threading.Thread(target=self._worker, args=(i, device), daemon=True) for i in range(depth)]
        for t in self._threads:
            t.start()

    def _worker(self, index, device):
        stream = None
        init_error = None
        if device is not None:
            try:
                check(lib.halide_b200_set_device(device))
                stream = lib.halide_b200_stream_create()
                if not stream:
                    raise RuntimeError("halide_b200_stream_create failed")
                lib.halide_b200_set_stream(stream)
            except BaseException as e:  # surfaced on the first job
                init_error = e
        q = self._queues[index]
        while True:
            job = q.get()
            if job is None:
                break
            ticket, fn, args = job
            try:
                if init_error is not None:
                    raise init_error
                out = (True, fn(*args))
            except BaseException as e:
                out = (False, e)
            with self._cv:
                self._done[ticket] = out
                self._cv.notify_all()
        if stream:
            lib.halide_b200_set_stream(None)
            lib.halide_b200_stream_destroy(stream)

    def submit(self, fn, *args, slot=None):
        with self._cv:
            ticket = self._next
            self._next += 1
        self._queues[(ticket if slot is None else slot) % self.depth].put((ticket, fn, args))
        return ticket

    def result(self, ticket):
        with self._cv:
            while ticket not in self._done:
                self._cv.wait()
            ok, val = self._done.pop(ticket)
        if not ok:
            raise val
        return val

    def close(self):
        if self._threads is None:
            return
        for q in self._queues:
            q.put(None)
        for t in self._threads:
            t.join()
        self._threads = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
