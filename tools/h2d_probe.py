"""Measure host->device upload strategies for a pageable numpy matrix (run under gpurun)."""
import threading
import time

import numpy as np
import torch

n, G = 25000, 20000
X = np.random.default_rng(0).random((n, G), dtype=np.float32)
dev = torch.device("cuda:0")
dst = torch.empty((n, G), device=dev)
torch.cuda.synchronize()
gb = X.nbytes / 1e9


def t(fn, name):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name:40s} {dt*1e3:8.1f} ms  {gb/dt:6.1f} GB/s")


t(lambda: dst.copy_(torch.from_numpy(X)), "direct pageable copy_ (cold)")
t(lambda: dst.copy_(torch.from_numpy(X)), "direct pageable copy_")


def ring(chunk_rows, nbuf):
    bufs = [torch.empty((chunk_rows, G), dtype=torch.float32, pin_memory=True) for _ in range(nbuf)]
    evs = [None] * nbuf

    def run():
        i = 0
        for r0 in range(0, n, chunk_rows):
            r1 = min(n, r0 + chunk_rows)
            b = i % nbuf
            if evs[b] is not None:
                evs[b].synchronize()
            bufs[b][: r1 - r0].copy_(torch.from_numpy(X[r0:r1]))
            dst[r0:r1].copy_(bufs[b][: r1 - r0], non_blocking=True)
            e = torch.cuda.Event()
            e.record()
            evs[b] = e
            i += 1
    return run


t0 = time.perf_counter()
r = ring(800, 3)
print(f"pinned ring alloc (3 x 64 MB): {(time.perf_counter()-t0)*1e3:.1f} ms")
t(r, "pinned ring 3x64MB, torch host copy")
t(r, "pinned ring 3x64MB, torch host copy (2)")
r = ring(3200, 3)
t(r, "pinned ring 3x256MB")


def threaded(nthreads):
    def run():
        def work(k):
            s = torch.cuda.Stream()
            rows = (n + nthreads - 1) // nthreads
            r0, r1 = k * rows, min(n, (k + 1) * rows)
            with torch.cuda.stream(s):
                dst[r0:r1].copy_(torch.from_numpy(X[r0:r1]), non_blocking=True)
            s.synchronize()
        th = [threading.Thread(target=work, args=(k,)) for k in range(nthreads)]
        [x.start() for x in th]
        [x.join() for x in th]
    return run


for k in (2, 4, 8):
    t(threaded(k), f"{k} threads, pageable copies on own streams")
t0 = time.perf_counter()
rc = torch.cuda.cudart().cudaHostRegister(X.ctypes.data, X.nbytes, 0)
print(f"cudaHostRegister 2 GB: {(time.perf_counter()-t0)*1e3:.1f} ms rc={rc}")
t(lambda: dst.copy_(torch.from_numpy(X), non_blocking=True), "copy from registered memory")
torch.cuda.cudart().cudaHostUnregister(X.ctypes.data)
print("threads", torch.get_num_threads())
