"""Generate tests/golden/*.npz by running the UNMODIFIED reference on seeded inputs.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/gen_golden.py

For every case it drives the reference's own ``ALPINE._initialize_matrices``
(main.py:436-472), ``_fit`` with ``max_iter=1`` repeatedly (main.py:486-676,
state persists in the AlpineMatrices), ``_compute_loss`` (main.py:726-753, via
``loss_history``), ``_scale_matrices`` (main.py:772-781), the ``_transform``
update loop (main.py:705-709, restated inline because ``_transform`` itself
needs an AnnData) and the gene-score arithmetic of
``get_covariate_gene_scores`` (main.py:256-261), and stores inputs, initial
factors and outputs.  The fixtures travel to the GPU box; the reference does
not.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from alpine_b200.utils.synth import make_counts, make_labels, make_poisson_counts  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: dict(shape, model kwargs, extras)
    "kl_basic": dict(
        n_cells=160, n_genes=96, cats=[3], rank=5, n_iter=10,
        kw=dict(n_components=6, n_covariate_components=[3], lam=[1e3]),
    ),
    "kl_reg_nan": dict(
        n_cells=203, n_genes=130, cats=[3, 4], rank=6, n_iter=10, nan_fraction=0.1,
        kw=dict(n_components=9, n_covariate_components=[4, 3], lam=[1e3, 5e2],
                orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5),
    ),
    "frob_reg": dict(
        n_cells=203, n_genes=130, cats=[3, 4], rank=6, n_iter=10, nan_fraction=0.05,
        kw=dict(n_components=9, n_covariate_components=[4, 3], lam=[1e3, 5e2],
                orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5, loss_type="frobenius"),
    ),
    "kl_lam0": dict(
        n_cells=120, n_genes=70, cats=[2], rank=0, n_iter=6,
        kw=dict(n_components=5, n_covariate_components=[2], lam=[0.0], alpha_W=1.5, l1_ratio_W=1.0),
    ),
    "als_reg": dict(
        n_cells=140, n_genes=90, cats=[3, 2], rank=5, n_iter=6,
        kw=dict(n_components=6, n_covariate_components=[3, 2], lam=[1e2, 1e3],
                orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5, use_als=True),
    ),
    # mini-batch epochs (main.py:509-521): ragged last batch; index streams recorded from the reference's sampler
    "mb_random": dict(
        n_cells=203, n_genes=130, cats=[3, 4], rank=6, n_iter=6, nan_fraction=0.05, batch_size=64,
        kw=dict(n_components=9, n_covariate_components=[4, 3], lam=[1e3, 5e2],
                orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5),
    ),
    # class-balanced sampling with replacement (sampling.py:18-33): duplicate cells inside a batch
    "mb_weighted": dict(
        n_cells=160, n_genes=96, cats=[3], rank=5, n_iter=6, batch_size=50, sampling_method="weighted",
        kw=dict(n_components=6, n_covariate_components=[3], lam=[1e3], alpha_W=0.3),
    ),
    "mb_als": dict(
        n_cells=150, n_genes=90, cats=[3, 2], rank=5, n_iter=5, batch_size=64,
        kw=dict(n_components=6, n_covariate_components=[3, 2], lam=[1e2, 1e3],
                orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5, use_als=True),
    ),
    "kl_long200": dict(
        n_cells=500, n_genes=300, cats=[3], rank=8, n_iter=200, keep_every=200,
        kw=dict(n_components=9, n_covariate_components=[3], lam=[1e3],
                orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5),
    ),
    # raw counts (tf32-exact: the 2-MMA kernel variant), 2,000 genes: top-100 rankings of W columns and of the
    # per-category gene scores (main.py:246-273) after 60 iterations; X stored as uint16
    "kl_scores2k": dict(
        n_cells=500, n_genes=2000, cats=[3, 4], rank=8, n_iter=60, keep_every=60, counts=True,
        kw=dict(n_components=10, n_covariate_components=[4, 3], lam=[1e3, 5e2],
                orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5),
    ),
}


def run_case(ref_main, name: str, spec: dict) -> dict:
    n, G = spec["n_cells"], spec["n_genes"]
    if spec.get("counts"):
        Xcg = make_poisson_counts(n, G, seed=len(name) * 7 + n, rank=spec["rank"])
    else:
        Xcg = make_counts(n, G, seed=len(name) * 7 + n, rank=spec["rank"])
    labels = make_labels(n, spec["cats"], seed=G, nan_fraction=spec.get("nan_fraction", 0.0))
    keys = [f"cov{i}" for i in range(len(labels))]
    index = [str(i) for i in range(n)]
    obs = pd.DataFrame({k: pd.Series(l, dtype=object, index=index) for k, l in zip(keys, labels)}, index=index)

    X = np.asarray(Xcg).astype(np.float32).T  # main.py:104 (F-order view, genes x cells)
    from alpine.utils.encoder import FeatureEncoders  # the reference's own encoder

    fe = FeatureEncoders(keys)
    Y = fe.fit_transform(obs) if keys else []

    model = ref_shim.make_reference_model(ref_main, n, keys, max_iter=1, **spec["kw"])
    epoch_streams = []
    if "batch_size" in spec:
        model.batch_size = spec["batch_size"]
        model.sampling_method = spec.get("sampling_method", "random")
        sampler = ref_main.generate_epoch_indices  # the reference's own sampler; only its output is recorded

        def recording_sampler(*a, **k):
            idx = sampler(*a, **k)
            epoch_streams.append(idx.cpu().numpy().astype(np.int64))
            return idx

        ref_main.generate_epoch_indices = recording_sampler
        torch.manual_seed(spec.get("sampler_seed", 7))
    mats = model._initialize_matrices(X, Y)
    out = {
        "X_cells_by_genes": Xcg,
        "n_cov": np.int64(len(keys)),
        "blocks": np.asarray(model.n_all_components, dtype=np.int64),
        "W0": torch.cat(mats.Ws, 1).numpy().copy(),
        "H0": torch.cat(mats.Hs, 0).numpy().copy(),
    }
    for i, y in enumerate(Y):
        out[f"Y{i}_cells_by_cat"] = y
        out[f"B0_{i}"] = mats.Bs[i].numpy().copy()
        out[f"labels{i}"] = np.asarray(["" if (isinstance(v, float) and v != v) else v for v in labels[i]])
        out[f"labels{i}_isna"] = np.asarray([isinstance(v, float) and v != v for v in labels[i]])
        out[f"cats{i}"] = np.asarray(fe.encoded_labels[keys[i]])

    keep_every = spec.get("keep_every", 1)
    losses, kept = [], []
    for it in range(spec["n_iter"]):
        model._fit(mats)  # one reference iteration
        losses.append(model.loss_history.iloc[-1].to_numpy(dtype=np.float64))
        if (it + 1) % keep_every == 0:
            kept.append(it + 1)
            out[f"W_it{it + 1}"] = torch.cat(mats.Ws, 1).numpy().copy()
            out[f"H_it{it + 1}"] = torch.cat(mats.Hs, 0).numpy().copy()
            for i in range(len(keys)):
                out[f"B{i}_it{it + 1}"] = mats.Bs[i].numpy().copy()
    out["kept_iters"] = np.asarray(kept, dtype=np.int64)
    if "batch_size" in spec:
        ref_main.generate_epoch_indices = sampler
        out["batch_size"] = np.int64(spec["batch_size"])
        for it, idx in enumerate(epoch_streams):
            out[f"epoch_idx_it{it + 1}"] = idx
    out["loss_history_ref_fp32"] = np.asarray(losses)

    # fp64 re-evaluation of the reference's loss formula from its final factors
    Wf = torch.cat(mats.Ws, 1).double()
    Hf = torch.cat(mats.Hs, 0).double()
    out["final_recon_fp64"] = np.float64((torch.norm(mats.X.double() - Wf @ Hf) ** 2).item())

    # scaling (main.py:772-781)
    model._scale_matrices(mats)
    out["W_scaled"] = torch.cat(mats.Ws, 1).numpy().copy()
    out["H_scaled"] = torch.cat(mats.Hs, 0).numpy().copy()
    for i in range(len(keys)):
        out[f"B{i}_scaled"] = mats.Bs[i].numpy().copy()

    # gene scores (main.py:256-261) on the scaled matrices, as fit() stores them
    npm = mats.to_numpy()
    for i in range(len(keys)):
        W, H, Yi = npm["Ws"][i], npm["Hs"][i], npm["Ys"][i]
        HY = H @ Yi.T / Yi.sum(axis=1)
        out[f"gene_scores{i}"] = W @ HY

    # transform loop (main.py:705-709) with the scaled W and a seeded H0
    g = torch.Generator().manual_seed(1234)
    Ht = torch.rand((model.total_components, n), dtype=torch.float32, generator=g)
    out["Ht0"] = Ht.numpy().copy()
    W = torch.cat(mats.Ws, 1)
    for _ in range(5):
        numerator = 2 * W.T @ mats.X
        denominator = 2 * W.T @ (W @ Ht)
        denominator = torch.clamp(denominator, min=model.eps)
        Ht *= numerator / denominator
    out["Ht_5"] = Ht.numpy().copy()
    return out


def main() -> None:
    ref_main = ref_shim.import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # deterministic summation order in the fixtures
    only = set(sys.argv[1:])
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        data = run_case(ref_main, name, spec)
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **data)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, final loss {data['loss_history_ref_fp32'][-1][:2]}")


if __name__ == "__main__":
    main()
