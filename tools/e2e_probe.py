"""Where the time of ALPINE.fit on host buffers goes (cfg 3 shapes by default): phases and upload details of
consecutive fits in one process, from pageable and from page-locked host memory.

    python tools/e2e_probe.py [--cells 100000] [--genes 20000] [--iters 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import pandas as pd
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpine_b200 import ALPINE  # noqa: E402
from alpine_b200.utils.anndata_compat import AnnData  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    n, G = a.cells, a.genes
    rng = np.random.default_rng(0)
    X = torch.rand((n, G)).pow_(3).numpy()
    obs = pd.DataFrame({"cov0": pd.Series([f"a{v}" for v in rng.integers(0, 3, n)], dtype=object),
                        "cov1": pd.Series([f"b{v}" for v in rng.integers(0, 4, n)], dtype=object)})
    kw = dict(n_components=90, n_covariate_components=[5, 5], lam=[1e3, 1e3], orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5)
    for label in ("pageable #1", "pageable #2", "pinned #1", "pinned #2"):
        if label == "pinned #1":
            t0 = time.perf_counter()
            rc = torch.cuda.cudart().cudaHostRegister(X.ctypes.data, X.nbytes, 0)
            print(f"cudaHostRegister of {X.nbytes / 1e9:.1f} GB: rc={rc}, {time.perf_counter() - t0:.2f} s")
        ad = AnnData(X, obs=obs.copy())
        model = ALPINE(device="cuda:0", **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.fit(ad, ["cov0", "cov1"], max_iter=a.iters)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{label}: fit {dt:.3f} s ({a.iters / dt:.1f} it/s)  phases {json.dumps({k: round(v, 3) for k, v in model.timings.items()})}"
              f"  detail {json.dumps({k: round(v, 3) for k, v in model.timings_detail.items()})}")


if __name__ == "__main__":
    main()
