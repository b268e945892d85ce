"""Cell-sharded fit == single-GPU fit, checked with the library's own kernels (no oracle involved).

Called under torchrun (one process per GPU, NCCL initialised) by ``tools/dist_check.py``, by the multi-GPU parity
test of ``tests/test_gpu_multi.py`` and, once and outside every timed region, by ``bench.py``'s N > 1 arm
(``"parity_vs_n1"`` in its JSON line).  Every rank runs the sharded fit twice -- exchanging through the NCCL
all-reduce of ``MUEngine.step`` and through the NVLink peer-memory kernels of ``csrc/peer_exchange.cuh`` -- and then
the same problem unsharded on its own GPU; it compares W, B, its column block of H and the loss history.  SURVEY.md
8 e1: N-GPU != 1-GPU bitwise (summation order), so the bar is 1e-5 Frobenius-relative; W must be bit-identical
across the ranks of one run.
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch


def _rel(a: np.ndarray, b: np.ndarray) -> float:
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b.astype(np.float64)), 1e-300))


def sharded_vs_single(dev: torch.device, n: int = 6001, G: int = 1500, blocks: Sequence[int] = (5, 5, 90),
                      cats: Sequence[int] = (3, 4), n_iter: int = 8, seed: int = 0, peer: bool = True) -> Dict:
    """Collective over the default process group.  Returns the worst errors over ranks (identical on every rank)."""
    import torch.distributed as dist

    from .. import _native
    from ..engine import MUEngine, shard_bounds

    rank, world = dist.get_rank(), dist.get_world_size()
    blocks, cats = list(blocks), list(cats)
    kw = dict(lam=[1e3, 5e2][: len(cats)], alpha_W=0.5, l1_ratio_W=0.5, orth_W=0.2)
    K = sum(blocks)
    g = torch.Generator(device="cpu").manual_seed(seed)  # the same host data on every rank
    Xh = torch.rand((n, G), generator=g).pow(3.0)
    Wh = torch.rand((G, K), generator=g).clamp(min=1e-6)
    Hh = torch.rand((K, n), generator=g).clamp(min=1e-6)
    Bh = [torch.rand((c, k), generator=g).clamp(min=1e-6) for c, k in zip(cats, blocks)]
    Yh = [torch.nn.functional.one_hot(torch.randint(0, c, (n,), generator=g), c).T.float().contiguous() for c in cats]

    def run(lo, hi, sharded, use_peer=False):
        X = _native.padded_rows(hi - lo, G, dev)
        X.copy_(Xh[lo:hi])
        H = _native.padded_rows(K, hi - lo, dev)
        H.copy_(Hh[:, lo:hi])
        W = Wh.clone().to(dev)
        Bs = [b.clone().to(dev) for b in Bh]
        Ys = [y[:, lo:hi].contiguous().to(dev) for y in Yh]
        s = _native.Solver(dev, G, hi - lo, blocks, cats)
        s.bind_dense(X)
        s.bind_labels(Ys)
        s.bind_factors(W, H, Bs)
        s.set_hparams(kw["lam"], kw["alpha_W"], kw["l1_ratio_W"], kw["orth_W"], 1e-6)
        peer_on = bool(use_peer and s.enable_peer_exchange())
        eng = MUEngine(s, kw["lam"])
        if not sharded:
            eng.world = 1
        hist = eng.run(n_iter)
        out = (W.cpu().numpy(), H.cpu().numpy(), [b.cpu().numpy() for b in Bs], hist, peer_on)
        s.close(collective=sharded)
        return out

    lo, hi = shard_bounds(n, world, rank)
    res = {"world": world, "shape": {"n_cells": n, "n_genes": G, "K": K, "n_iter": n_iter,
                                     "cells_per_rank": [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0]
                                                        for r in range(world)]}}
    W1, H1, B1, hist1, _ = run(0, n, False)
    modes = [("nccl", False)] + ([("peer", True)] if peer else [])
    ok = True
    for name, use_peer in modes:
        Wd, Hd, Bd, hist, peer_on = run(lo, hi, True, use_peer)
        gathered = [None] * world
        dist.all_gather_object(gathered, Wd.tobytes())
        same_w = all(x == gathered[0] for x in gathered)
        errs = np.array([_rel(Wd, W1), _rel(Hd, H1[:, lo:hi]), max(_rel(a, b) for a, b in zip(Bd, B1)),
                         float(np.max(np.abs(hist[:, 1] - hist1[:, 1]) / np.abs(hist1[:, 1]))),
                         # the KL prediction terms cancel element-wise: compare them absolutely, per cell
                         float(np.max(np.abs(hist[:, 2:] - hist1[:, 2:]))) / n if hist.shape[1] > 2 else 0.0])
        t = torch.from_numpy(errs).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        eW, eH, eB, eL, eP = [float(v) for v in t.cpu()]
        res[name] = {"W": eW, "H": eH, "B": eB, "recon_loss": eL, "pred_loss_abs_per_cell": eP,
                     "W_bit_identical_across_ranks": bool(same_w), "peer_exchange_active": bool(peer_on)}
        ok = ok and max(eW, eH, eB, eL) < 1e-5 and eP < 1e-6 and same_w and (peer_on or not use_peer)
    res["ok"] = bool(ok)
    return res


def minibatch_sharded_vs_single(dev: torch.device, n: int = 3001, G: int = 700, batch_size: int = 512, epochs: int = 3,
                                sampling_method: str = "random", use_als: bool = False) -> Dict:
    """Mini-batch ``ALPINE.fit`` under cell sharding == the same fit on one GPU (same epoch index stream: every rank
    draws it from the same seeded generators).  Collective over the default process group."""
    import pandas as pd
    import torch.distributed as dist

    from .. import engine as engine_mod
    from .. import main as main_mod
    from ..utils.anndata_compat import AnnData
    from ..utils.synth import make_counts, make_labels

    X = make_counts(n, G, seed=3, rank=6)
    labels = make_labels(n, [3, 2], seed=3, nan_fraction=0.02)
    obs = pd.DataFrame({f"cov{i}": pd.Series(l, dtype=object) for i, l in enumerate(labels)})
    kw = dict(n_components=8, n_covariate_components=[3, 2], lam=[1e2, 5e1], orth_W=0.1, alpha_W=0.3, l1_ratio_W=0.5,
              use_als=use_als, device=str(dev), random_state=7)

    def fit():
        model = main_mod.ALPINE(**kw)
        model.fit(AnnData(X.copy(), obs=obs.copy()), ["cov0", "cov1"], batch_size=batch_size, max_iter=epochs,
                  sampling_method=sampling_method)
        mats = model.get_decomposed_matrices()
        return (np.concatenate(mats["Ws"], axis=1), np.concatenate(mats["Hs"], axis=0), mats["Bs"],
                model.loss_history.to_numpy())

    Wd, Hd, Bd, hist_d = fit()
    real_main, real_eng = main_mod.dist_info, engine_mod.dist_info
    main_mod.dist_info = engine_mod.dist_info = lambda group=None: (0, 1)  # the same fit, unsharded, on this GPU
    try:
        W1, H1, B1, hist_1 = fit()
    finally:
        main_mod.dist_info, engine_mod.dist_info = real_main, real_eng
    errs = np.array([_rel(Wd, W1), _rel(Hd, H1), max(_rel(a, b) for a, b in zip(Bd, B1)),
                     float(np.max(np.abs(hist_d[:, 1] - hist_1[:, 1]) / np.abs(hist_1[:, 1])))])
    t = torch.from_numpy(errs).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    eW, eH, eB, eL = [float(v) for v in t.cpu()]
    return {"world": dist.get_world_size(), "sampling_method": sampling_method, "use_als": use_als, "W": eW, "H": eH,
            "B": eB, "recon_loss": eL, "ok": bool(max(eW, eH, eB, eL) < 2e-5)}
