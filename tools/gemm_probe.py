"""GPU probe: time and accuracy of the two contractions on synthetic data (not a test; run under gpurun)."""
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
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check-rows", type=int, default=256)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    G, n, K = a.genes, a.cells, a.k
    torch.manual_seed(0)
    X = _native.padded_rows(n, G, dev)
    # Gamma(0.3, 2)-like positive data, generated on the device in row chunks
    step = max(1, (1 << 28) // max(G, 1))
    for r0 in range(0, n, step):
        r1 = min(n, r0 + step)
        X[r0:r1] = torch.distributions.Gamma(0.3, 0.5).sample((r1 - r0, G)).to(dev) if False else \
            torch.rand((r1 - r0, G), device=dev).pow_(3.0).mul_(4.0)
    W = torch.rand((G, K), device=dev)
    H = _native.padded_rows(K, n, dev)
    H.copy_(torch.rand((K, n), device=dev))
    s = _native.Solver(dev, G, n, [K], [])
    s.bind_dense(X)
    s.bind_factors(W, H, [])
    for name, fn, flops_bytes in (("xh", s.xh_product, None), ("wx", s.wx_product, None)):
        out = fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = min(ts)
        gbs = 4.0 * G * n / ms / 1e6
        tf = 2.0 * G * n * K / ms / 1e9
        print(f"{name}: best {ms:.3f} ms  median {sorted(ts)[len(ts)//2]:.3f} ms  X-stream {gbs:.0f} GB/s  "
              f"{tf:.1f} fp32-equiv TFLOP/s ({2*tf:.1f} tf32-MMA equivalents: tf32 hi*hi + two bf16 corrections)  {s.query()}")
        # accuracy on a subset against fp64
        r = a.check_rows
        if name == "xh":
            ref = (H.double() @ X[:, :r].double())            # (K, r)
            got = out[:, :r].double()
        else:
            ref = (W.double().T @ X[:r, :].double().T)        # (K, r)
            got = out[:, :r].double()
        rel = ((got - ref).norm() / ref.norm()).item()
        bias = ((got - ref) / ref).mean().item()
        print(f"   accuracy vs fp64 on {r} rows: rel_fro {rel:.3e}  mean signed rel err {bias:.3e}")
    print("launches", _native.launch_count())


if __name__ == "__main__":
    main()
