"""Shared helpers for the parity tests (fixtures -> oracle State / HyperParams)."""
from __future__ import annotations

import glob
import os

import numpy as np

from oracle import alpine_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# model kwargs of every golden case (must mirror oracle/gen_golden.py CASES)
CASE_KW = {
    "kl_basic": dict(n_components=6, n_covariate_components=[3], lam=[1e3]),
    "kl_reg_nan": dict(n_components=9, n_covariate_components=[4, 3], lam=[1e3, 5e2],
                       orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5),
    "frob_reg": dict(n_components=9, n_covariate_components=[4, 3], lam=[1e3, 5e2],
                     orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5, loss_type="frobenius"),
    "kl_lam0": dict(n_components=5, n_covariate_components=[2], lam=[0.0], alpha_W=1.5, l1_ratio_W=1.0),
    "als_reg": dict(n_components=6, n_covariate_components=[3, 2], lam=[1e2, 1e3],
                    orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5, use_als=True),
    "mb_random": dict(n_components=9, n_covariate_components=[4, 3], lam=[1e3, 5e2],
                      orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5),
    "mb_weighted": dict(n_components=6, n_covariate_components=[3], lam=[1e3], alpha_W=0.3),
    "mb_als": dict(n_components=6, n_covariate_components=[3, 2], lam=[1e2, 1e3],
                   orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5, use_als=True),
    "kl_scores2k": dict(n_components=10, n_covariate_components=[4, 3], lam=[1e3, 5e2],
                        orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5),
    "kl_long200": dict(n_components=9, n_covariate_components=[3], lam=[1e3],
                       orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5),
}


MINIBATCH_CASES = ("mb_random", "mb_weighted", "mb_als")


def full_batch_mu_names():
    """Fixtures of the full-batch, non-ALS loop (what alpine_mu_partials / alpine_mu_apply iterate)."""
    return [n for n in golden_names() if not CASE_KW[n].get("use_als", False) and n not in MINIBATCH_CASES]


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False))


def hp_of(name) -> orc.HyperParams:
    kw = dict(CASE_KW[name])
    kw.pop("use_als", None)
    return orc.HyperParams(**kw)


def inputs_of(g):
    """X (genes x cells, the reference's F-order view), Ys (c_i x n), initial State."""
    n_cov = int(g["n_cov"])
    X = np.asarray(g["X_cells_by_genes"]).astype(np.float32).T
    Ys = [np.ascontiguousarray(g[f"Y{i}_cells_by_cat"].T) for i in range(n_cov)]
    st = orc.State(g["W0"].copy(), g["H0"].copy(), [g[f"B0_{i}"].copy() for i in range(n_cov)],
                   [int(b) for b in g["blocks"]])
    return X, Ys, st


def epoch_batches(g, it):
    """Batch index vectors of epoch ``it`` (1-based): the recorded sampler stream cut as sampling.py:58-71 does;
    one ``None`` (= full batch in natural order) for the full-batch fixtures."""
    if "batch_size" not in g:
        return [None]
    idx, bs = g[f"epoch_idx_it{it}"], int(g["batch_size"])
    return [idx[b0:b0 + bs] for b0 in range(0, len(idx), bs)]


def rel_fro(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def assert_same_top_ranking(got, ref, top=100, tie_rel=1e-5, what=""):
    """Identical top-`top` ordering of two score vectors, except where the REFERENCE itself cannot tell two entries
    apart: a position may hold a different index only if that index's reference score is within `tie_rel` (relative)
    of the reference's own entry at that position.  fp32 implementations that differ in summation order agree to a
    few 1e-6; the reference's trajectories contain adjacent scores closer than that (kl_long200, component 2, ranks
    59 / 60: relative gap 8.6e-7), whose order no re-implementation can be held to."""
    got, ref = np.asarray(got), np.asarray(ref)
    a = np.argsort(-got, kind="stable")[:top]
    b = np.argsort(-ref, kind="stable")[:top]
    for pos in np.nonzero(a != b)[0]:
        ra, rb = float(ref[a[pos]]), float(ref[b[pos]])
        assert abs(ra - rb) <= tie_rel * abs(rb), (what, int(pos), int(a[pos]), int(b[pos]), ra, rb)
