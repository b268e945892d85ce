"""Trial-level scheduling: one fit per GPU.

``ComponentOptimizer.calc_score`` (reference optimization.py:220-287) runs ``n_splits`` independent fold fits per
trial, one after the other on a single device.  Folds (and, with batched suggestions, trials) are independent, so
here they are dispatched over the GPUs of the box: one worker thread per device, each pulling the next job from a
shared queue and running it with its own ``device="cuda:<i>"``.  Threads (not processes) are used because the jobs
share the large, read-only AnnData in host memory, ctypes releases the GIL inside the native calls, and every
native context selects its own device.  Results come back in job order; the first exception is re-raised after all
workers have stopped.
"""
from __future__ import annotations

import queue
import threading
import time
from typing import Any, Callable, List, Optional, Sequence


def visible_devices(device: str = "cuda") -> List[str]:
    """``["cuda:0", ...]`` for ``device="cuda"``; a single entry when an index was given or no GPU is present."""
    try:
        import torch

        if device == "cuda" and torch.cuda.is_available():
            return [f"cuda:{i}" for i in range(torch.cuda.device_count())]
    except Exception:
        pass
    return [device]


def _pool_limits(n_workers: int):
    """Context manager capping the native thread pools at cores / n_workers (no-op without threadpoolctl)."""
    import os
    from contextlib import nullcontext

    try:
        from threadpoolctl import threadpool_limits
    except ImportError:  # pragma: no cover
        return nullcontext()
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        cores = os.cpu_count() or 1
    return threadpool_limits(limits=max(1, cores // max(1, n_workers)))


class DeviceScheduler:
    """Run ``fn(job, device)`` for every job, at most one job per device at a time."""

    def __init__(self, devices: Sequence[str]):
        if not devices:
            raise ValueError("at least one device is required")
        self.devices = list(devices)
        self.assignments: List[tuple] = []  # (job index, device) in completion order, for inspection / tests
        self.busy_s = {d: 0.0 for d in self.devices}  # seconds every device spent inside jobs (bench.py's cfg5 block)

    def map(self, fn: Callable[[Any, str], Any], jobs: Sequence[Any]) -> List[Any]:
        results: List[Any] = [None] * len(jobs)
        errors: List[BaseException] = []
        todo: "queue.Queue[int]" = queue.Queue()
        for i in range(len(jobs)):
            todo.put(i)
        lock = threading.Lock()

        def worker(device: str) -> None:
            while not errors:
                try:
                    i = todo.get_nowait()
                except queue.Empty:
                    return
                try:
                    t0 = time.perf_counter()
                    out = fn(jobs[i], device)
                    dt = time.perf_counter() - t0
                    with lock:
                        results[i] = out
                        self.assignments.append((i, device))
                        self.busy_s[device] += dt
                except BaseException as exc:  # propagate to the caller
                    with lock:
                        errors.append(exc)
                    return

        n_workers = min(len(self.devices), max(1, len(jobs)))
        if n_workers == 1:
            worker(self.devices[0])
        else:
            threads = [threading.Thread(target=worker, args=(d,), daemon=True) for d in self.devices[:n_workers]]
            # the host-side parts of concurrent jobs (BLAS / OpenMP pools of numpy and scikit-learn inside the scorer)
            # would each start one thread per core: give every worker its share of the cores instead
            with _pool_limits(n_workers):
                for t in threads:
                    t.start()
                for t in threads:
                    t.join()
        if errors:
            raise errors[0]
        return results
