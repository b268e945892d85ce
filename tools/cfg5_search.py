"""BASELINE configs[4]: ComponentOptimizer.search_hyperparams on 5,000 HVG x 50,000 cells, one fit per GPU.

    python tools/cfg5_search.py --trials 64            # uses every visible GPU of the box (threads, one per GPU)

hyperopt / scanpy / leidenalg are not installed in this image, so the search is the seeded random search over the
reference's space and the score is ARI + homogeneity of a k-means clustering of the embedding (optimization.py
stand-ins); what this measures is the fold scheduler driving real fits + transforms on every GPU.
"""
import argparse
import os
import sys
import time
from collections import Counter

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpine_b200 import ComponentOptimizer  # noqa: E402
from alpine_b200.utils.anndata_compat import AnnData  # noqa: E402
from alpine_b200.utils.synth import make_counts, make_labels  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=64)
    ap.add_argument("--cells", type=int, default=50000)
    ap.add_argument("--genes", type=int, default=5000)
    ap.add_argument("--max-iter", type=int, default=100)
    ap.add_argument("--gpus", type=int, default=0, help="limit the number of GPUs used (0 = all visible)")
    a = ap.parse_args()
    X = make_counts(a.cells, a.genes, seed=0, rank=12)
    labels = make_labels(a.cells, [3, 4], seed=0)
    obs = pd.DataFrame({f"cov{i}": pd.Series(l, dtype=object) for i, l in enumerate(labels)})
    adata = AnnData(X, obs=obs)
    opt = ComponentOptimizer(adata, ["cov0", "cov1"], max_iter=a.max_iter, device="cuda", random_state=42)
    if a.gpus:
        opt.devices = opt.devices[: a.gpus]
    t0 = time.perf_counter()
    best = opt.search_hyperparams(n_total_components_range=(10, 100), n_splits=3, max_evals=a.trials)
    dt = time.perf_counter() - t0
    hist = opt.get_train_history()
    print(f"cfg5: {a.trials} trials x 3 folds on {a.genes} genes x {a.cells} cells, max_iter={a.max_iter}, "
          f"{len(opt.devices)} GPU(s): {dt:.1f} s  ({dt / a.trials:.2f} s per trial, {len(hist)} feasible trials)")
    print("  last batch's fold -> device assignments:", dict(Counter(d for _, d in opt.last_scheduler.assignments)))
    print("  best:", {k: (v if not isinstance(v, float) else round(v, 3)) for k, v in best.items()})
    print("  top scores:", [round(float(s), 4) for s in hist["score"].head(5)])


if __name__ == "__main__":
    main()
