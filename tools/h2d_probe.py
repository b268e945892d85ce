"""Host -> device upload rate of the library's uploader (csrc/host_upload.cuh) on this box.

    python tools/h2d_probe.py [--gb 8] [--threads 16]

Prints the first call (includes pinning the staging ring) and a second call, against a pinned-memory cudaMemcpy of the
same size (the PCIe ceiling) and torch's pageable copy (what the reference's torch.tensor(..., device=) does).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpine_b200 import _native  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=8.0)
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--threads", default="4,8,12,16")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    rows = int(a.gb * 1e9 / 4 / a.genes)
    src = np.ones((rows, a.genes), dtype=np.float32)
    src[::1000] = 2.0
    dst = _native.padded_rows(rows, a.genes, dev)
    torch.cuda.synchronize()
    gb = src.nbytes / 1e9
    for th in [int(v) for v in a.threads.split(",")] * 2:
        t0 = time.perf_counter()
        _native.upload_rows(dst, src, threads=th)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"uploader threads={th:2d}: {gb / (t2 - t0):6.1f} GB/s  (call returned after {t1 - t0:.3f} s, done after {t2 - t0:.3f} s)")
    assert float(dst[::1000].sum()) == 2.0 * a.genes * len(range(0, rows, 1000))
    t0 = time.perf_counter()
    pin = torch.empty((1 << 26,), dtype=torch.float32, pin_memory=True)
    print(f"pinning 256 MB: {time.perf_counter() - t0:.3f} s")
    flat = dst.reshape(-1) if dst.is_contiguous() else dst
    t0 = time.perf_counter()
    for i in range(8):
        flat.view(-1)[i * (1 << 26):(i + 1) * (1 << 26)].copy_(pin, non_blocking=True)
    torch.cuda.synchronize()
    print(f"pinned cudaMemcpy 2 GB: {2.147 / (time.perf_counter() - t0):.1f} GB/s")
    t0 = time.perf_counter()
    dst[: rows // 8].copy_(torch.from_numpy(src[: rows // 8]))
    torch.cuda.synchronize()
    print(f"torch pageable copy: {gb / 8 / (time.perf_counter() - t0):.1f} GB/s")


if __name__ == "__main__":
    main()
