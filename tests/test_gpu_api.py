"""GPU tests of the drop-in Python API (ALPINE.fit / transform / scores) against the oracle."""
import numpy as np
import pandas as pd
import pytest
import torch

from alpine_b200 import ALPINE
from alpine_b200.utils.anndata_compat import AnnData
from alpine_b200.utils.synth import make_counts, make_labels
from oracle import alpine_oracle as orc
from tests.helpers import rel_fro

pytestmark = pytest.mark.gpu


def _adata(n=600, G=400, cats=(3, 4), seed=1, nan_fraction=0.05):
    X = make_counts(n, G, seed=seed, rank=6)
    labels = make_labels(n, list(cats), seed=seed, nan_fraction=nan_fraction)
    obs = pd.DataFrame({f"cov{i}": pd.Series(l, dtype=object) for i, l in enumerate(labels)})
    obs.index = [f"cell{i}" for i in range(n)]
    var = pd.DataFrame(index=[f"gene{i}" for i in range(G)])
    return AnnData(X, obs=obs, var=var)


KW = dict(n_components=8, n_covariate_components=[3, 4], lam=[1e3, 5e2], orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5)


@pytest.mark.parametrize("loss_type", ["kl-divergence", "frobenius"])
def test_fit_matches_oracle_from_same_initialisation(loss_type):
    ad = _adata()
    keys = ["cov0", "cov1"]
    n_iter = 10
    model = ALPINE(device="cuda:0", loss_type=loss_type, **KW)
    # replicate fit()'s preparation, snapshot the initial factors, then run the hot loop
    model.covariate_keys, model.sampling_method, model.verbose = keys, "random", False
    from alpine_b200.utils.encoder import FeatureEncoders

    model.fe = FeatureEncoders(keys)
    Y = model.fe.fit_transform(ad.obs)
    X = np.ascontiguousarray(ad.X, dtype=np.float32).T
    model.batch_size, model.max_iter = X.shape[1], n_iter
    m = model._initialize_matrices(X, Y)
    st = orc.State(m.W.cpu().numpy().copy(), m.H.cpu().numpy().copy(), [b.cpu().numpy().copy() for b in m.Bs],
                   list(model.n_all_components))
    hp = orc.HyperParams(loss_type=loss_type, **KW)
    Ys = [np.ascontiguousarray(y.T) for y in Y]
    model._fit(m)
    hist_ref, _ = orc.fit_loop(X, Ys, st, hp, n_iter)
    assert rel_fro(m.W.cpu().numpy(), st.W) < 2e-5
    assert rel_fro(m.H.cpu().numpy(), st.H) < 2e-5
    for b, bo in zip(m.Bs, st.Bs):
        assert rel_fro(b.cpu().numpy(), bo) < 2e-5
    lh = model.loss_history
    assert list(lh.columns) == ["total loss", "reconstruction loss", "prediction loss(cov0)", "prediction loss(cov1)"]
    assert len(lh) == n_iter
    ref64 = orc.compute_loss(X, Ys, st, hp, dtype=np.float64)
    assert abs(lh["reconstruction loss"].iloc[-1] - ref64[1]) / ref64[1] < 1e-4
    np.testing.assert_allclose(lh.iloc[-1, 2:].to_numpy(dtype=float), ref64[2:], rtol=1e-3, atol=1e-6 * X.shape[1])
    np.testing.assert_allclose(lh["total loss"].iloc[-1], ref64[0], rtol=1e-4)
    # _compute_loss evaluates the same quantities from the factors as they are
    now = model._compute_loss(m)
    np.testing.assert_allclose(now[1], ref64[1], rtol=1e-4)
    np.testing.assert_allclose(now[2:], ref64[2:], rtol=1e-3, atol=1e-6 * X.shape[1])
    # scaling (main.py:772-781)
    model._scale_matrices(m)
    orc.scale_matrices(st, hp)
    assert rel_fro(m.W.cpu().numpy(), st.W) < 2e-5
    assert rel_fro(m.H.cpu().numpy(), st.H) < 2e-5


def test_public_fit_transform_scores_roundtrip():
    ad = _adata()
    keys = ["cov0", "cov1"]
    model = ALPINE(device="cuda", **KW)
    out = model.fit(ad, keys, max_iter=15)
    assert out is model
    K = model.total_components
    mats = model.get_decomposed_matrices()
    assert set(mats) == {"X", "Ys", "Ws", "Hs", "Bs"}
    assert [w.shape for w in mats["Ws"]] == [(400, 3), (400, 4), (400, 8)]
    assert [h.shape for h in mats["Hs"]] == [(3, 600), (4, 600), (8, 600)]
    assert mats["X"].shape == (400, 600)
    for w in mats["Ws"]:  # scale_needed: every column sums to one
        np.testing.assert_allclose(w.sum(axis=0), 1.0, rtol=1e-5)
    assert ad.obsm["ALPINE_embedding"].shape == (600, 8) and ad.varm["ALPINE_weights"].shape == (400, 8)
    assert ad.obsm["cov0"].shape == (600, 3) and ad.obsm["cov0_dummy_matrix"].shape[0] == 600
    # the objective decreased
    lh = model.loss_history["total loss"].to_numpy()
    assert lh[-1] < lh[0]
    # gene scores (main.py:246-273) vs the oracle arithmetic on the same matrices
    scores = model.get_covariate_gene_scores()
    ref = orc.covariate_gene_scores(mats["Ws"][:2], mats["Hs"][:2], mats["Ys"])
    for key, r in zip(keys, ref):
        assert list(scores[key].index) == model.feature_names
        np.testing.assert_allclose(scores[key].to_numpy(), r, rtol=1e-5)
    assert model.get_covariate_gene_scores(ad) is None and "cov0_gene_scores" in ad.varm
    # transform on held-out cells: same H as the oracle loop from the same H0
    val = _adata(n=200, seed=9)
    torch.manual_seed(7)
    torch.cuda.manual_seed(7)
    state = torch.cuda.get_rng_state(0)
    model.transform(val, n_iter=6)
    torch.cuda.set_rng_state(state, 0)
    H0 = torch.rand((K, 200), dtype=torch.float32, device="cuda:0").cpu().numpy()
    W = np.concatenate(mats["Ws"], axis=1)
    Ht = orc.transform_loop(np.ascontiguousarray(val.X).T, W, H0, 6, model.eps)
    got = np.concatenate([val.obsm["cov0"].T, val.obsm["cov1"].T, val.obsm["ALPINE_embedding"].T], axis=0)
    assert rel_fro(got, Ht) < 2e-5
    assert model.transform(val) is None  # n_iter defaults to max_iter
    total = model.compute_loss(val)
    assert np.isfinite(total) and total > 0
    model.get_normalized_expression(val, library_size=1e4)
    np.testing.assert_allclose(val.layers["normalized_expression"].sum(axis=1), 1e4, rtol=1e-4)


def test_fit_with_automatic_max_iter():
    ad = _adata(n=300, G=200, cats=(3,), nan_fraction=0.0)
    model = ALPINE(n_components=6, n_covariate_components=[3], lam=[1e2], device="cuda")
    with pytest.warns(None) if False else __import__("contextlib").nullcontext():
        model.fit(ad, ["cov0"])  # warm-up of 200 iterations + Kneedle elbow (main.py:116-131)
    assert 1 <= model.max_iter <= 200
    assert len(model.loss_history) == model.max_iter


def test_als_public_fit_matches_oracle():
    ad = _adata(n=400, G=250, cats=(3, 4), nan_fraction=0.03)
    keys = ["cov0", "cov1"]
    model = ALPINE(device="cuda:0", use_als=True, **KW)
    model.covariate_keys, model.sampling_method, model.verbose = keys, "random", False
    from alpine_b200.utils.encoder import FeatureEncoders

    model.fe = FeatureEncoders(keys)
    Y = model.fe.fit_transform(ad.obs)
    X = np.ascontiguousarray(ad.X, dtype=np.float32).T
    model.batch_size, model.max_iter = X.shape[1], 6
    m = model._initialize_matrices(X, Y)
    st = orc.State(m.W.cpu().numpy().copy(), m.H.cpu().numpy().copy(), [b.cpu().numpy().copy() for b in m.Bs],
                   list(model.n_all_components))
    hp = orc.HyperParams(**KW)
    Ys = [np.ascontiguousarray(y.T) for y in Y]
    model._fit(m)
    orc.fit_loop(X, Ys, st, hp, 6, use_als=True)
    assert rel_fro(m.W.cpu().numpy(), st.W) < 2e-5 and rel_fro(m.H.cpu().numpy(), st.H) < 2e-5
    ref64 = orc.compute_loss(X, Ys, st, hp, dtype=np.float64)
    assert abs(model.loss_history["reconstruction loss"].iloc[-1] - ref64[1]) / ref64[1] < 1e-4
    out = ALPINE(device="cuda", use_als=True, **KW).fit(ad, keys, max_iter=5)
    assert len(out.loss_history) == 5 and np.isfinite(out.loss_history.to_numpy()).all()


@pytest.mark.parametrize("method", ["random", "weighted"])
def test_sparse_minibatch_equals_dense_minibatch(method):
    """Mini-batch epochs on CSR input: the batch's rows are re-tiled on the device; same index stream => same fit."""
    import scipy.sparse as sp

    ad = _adata(n=500, G=300, cats=(3,), nan_fraction=0.02)
    ad.X[np.random.default_rng(3).random(ad.X.shape) < 0.6] = 0.0
    ad_sp = AnnData(sp.csr_matrix(ad.X), obs=ad.obs.copy(), var=ad.var.copy())
    rng = np.random.default_rng(9)
    streams = [rng.permutation(500) if method == "random" else rng.integers(0, 500, 500) for _ in range(3)]
    fits = []
    for data in (ad, ad_sp):
        model = ALPINE(n_components=5, n_covariate_components=[3], lam=[1e2], alpha_W=0.1, device="cuda")
        model._epoch_index_stream = iter([s.copy() for s in streams])
        model.fit(data, ["cov0"], max_iter=3, batch_size=128, sampling_method=method)
        fits.append((data.obsm["ALPINE_embedding"].copy(), data.varm["ALPINE_weights"].copy(),
                     model.loss_history.to_numpy()))
    np.testing.assert_allclose(fits[1][0], fits[0][0], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(fits[1][1], fits[0][1], rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(fits[1][2], fits[0][2], rtol=1e-5)


def test_unsupported_modes_raise():
    ad = _adata(n=200, G=100, cats=(3,), nan_fraction=0.0)
    with pytest.raises(ValueError):
        ALPINE(n_components=4, n_covariate_components=[3], lam=[1.0]).fit(ad, ["cov0"], max_iter=2, sampling_method="nope")


def _model_on_golden(name, g, **extra):
    """ALPINE + AlpineMatrices holding a golden fixture's inputs and initial factors on cuda:0."""
    from alpine_b200 import _native
    from alpine_b200.main import AlpineMatrices
    from tests.gpu_utils import to_dev_padded
    from tests.helpers import CASE_KW

    kw = dict(CASE_KW[name])
    kw.update(extra)
    dev = torch.device("cuda:0")
    n_cov = int(g["n_cov"])
    model = ALPINE(device="cuda:0", **kw)
    model.covariate_keys = [f"cov{i}" for i in range(n_cov)]
    model.verbose = False
    Xd = to_dev_padded(g["X_cells_by_genes"], dev)
    n = Xd.shape[0]
    W = torch.from_numpy(g["W0"].copy()).to(dev)
    H = to_dev_padded(g["H0"], dev)
    Ys = [torch.from_numpy(np.ascontiguousarray(g[f"Y{i}_cells_by_cat"].T)).to(dev) for i in range(n_cov)]
    Bs = [torch.from_numpy(g[f"B0_{i}"].copy()).to(dev) for i in range(n_cov)]
    Ws, Hs, col = [], [], 0
    for k in model.n_all_components:
        Ws.append(W[:, col:col + k])
        Hs.append(H[col:col + k, :])
        col += k
    m = AlpineMatrices(X=Xd.T, Ys=Ys, Ws=Ws, Hs=Hs, Bs=Bs, W=W, H=H, X_cells_major=Xd, shard=(0, n), n_total=n)
    return model, m


@pytest.mark.parametrize("name", ["mb_random", "mb_weighted", "mb_als"])
def test_minibatch_epochs_match_reference_golden(name):
    """Mini-batch epochs (main.py:509-521, 589-663) replaying the index streams the reference's sampler produced
    (oracle/gen_golden.py): ragged last batch, and duplicate cells inside a batch for the weighted sampler."""
    from tests.helpers import load_golden

    g = load_golden(name)
    model, m = _model_on_golden(name, g)
    n_iter = int(g["kept_iters"][-1])
    model.batch_size = int(g["batch_size"])
    model.sampling_method = "weighted" if name == "mb_weighted" else "random"
    model.max_iter = 1
    ref_hist = g["loss_history_ref_fp32"]
    for it in range(1, n_iter + 1):
        model._epoch_index_stream = iter([g[f"epoch_idx_it{it}"]])
        model._fit(m)  # one epoch, as the fixture generator drives the reference
        assert rel_fro(m.W.cpu().numpy(), g[f"W_it{it}"]) < 1e-4, (name, it, "W")
        assert rel_fro(m.H.cpu().numpy(), g[f"H_it{it}"]) < 1e-4, (name, it, "H")
        for i, b in enumerate(m.Bs):
            assert rel_fro(b.cpu().numpy(), g[f"B{i}_it{it}"]) < 1e-4, (name, it, f"B{i}")
        row = model.loss_history.iloc[-1].to_numpy(dtype=float)
        np.testing.assert_allclose(row[1], ref_hist[it - 1][1], rtol=2e-3)
        np.testing.assert_allclose(row[2:], ref_hist[it - 1][2:], rtol=1e-3, atol=1e-6 * m.n_total)
    recon = model.loss_history["reconstruction loss"].iloc[-1]
    assert abs(recon - float(g["final_recon_fp64"])) / float(g["final_recon_fp64"]) < 1e-4


def test_minibatch_public_fit_runs_with_own_sampler():
    ad = _adata(n=500, G=300, cats=(3,), nan_fraction=0.0)
    for method in ("random", "weighted"):
        model = ALPINE(n_components=5, n_covariate_components=[3], lam=[1e2], device="cuda")
        model.fit(ad, ["cov0"], max_iter=4, batch_size=128, sampling_method=method)
        lh = model.loss_history["total loss"].to_numpy()
        assert len(lh) == 4 and np.all(np.isfinite(lh)) and lh[-1] < lh[0]


def test_large_negative_matrix_is_rejected_after_the_deferred_device_check():
    """The non-negativity scan of a large dense X runs on the device after the upload; the error is the reference's
    (main.py:399-400)."""
    n, G = 5000, 4000  # > 2**24 elements: the host scan is deferred
    X = np.ones((n, G), dtype=np.float32)
    X[4321, 1234] = -0.5
    obs = pd.DataFrame({"cov0": pd.Series(["a", "b"] * (n // 2), dtype=object)})
    ad = AnnData(X, obs=obs)
    model = ALPINE(n_components=4, n_covariate_components=[2], lam=[1.0], device="cuda")
    with pytest.raises(ValueError, match="non-negative"):
        model.fit(ad, ["cov0"], max_iter=2)
    X[4321, 1234] = np.nan
    with pytest.raises(ValueError, match="non-negative"):
        model.fit(ad, ["cov0"], max_iter=2)
    X[4321, 1234] = 0.5
    assert model.fit(ad, ["cov0"], max_iter=2) is model


# ------------------------------------------------------------------------------------------------- random streams
def _reference_draw_sequence(model, G, n, cats, dev):
    """The draws the reference makes on the process-wide generators, restated call by call: seeding
    (main.py:440-442), then rand for every W block, every H block, every B (main.py:454-470)."""
    torch.manual_seed(model.random_state)
    torch.cuda.manual_seed(model.random_state)
    for k in model.n_all_components:
        torch.rand((G, k), dtype=torch.float32, device=dev)
    for k in model.n_all_components:
        torch.rand((k, n), dtype=torch.float32, device=dev)
    for c, k in zip(cats, model.n_covariate_components):
        torch.rand((c, k), dtype=torch.float32, device=dev)


def test_generators_after_fit_are_where_the_reference_leaves_them():
    """fit() draws W / H / B and one randperm per full-batch iteration from the seeded device generator
    (main.py:440-470, 502-506); a following transform() continues that stream (main.py:687-689)."""
    dev = torch.device("cuda:0")
    ad = _adata()
    keys = ["cov0", "cov1"]
    n_iter = 7
    model = ALPINE(device="cuda:0", random_state=11, **KW)
    model.fit(ad, keys, max_iter=n_iter)
    got_state = torch.cuda.get_rng_state(0).clone()
    _reference_draw_sequence(model, 400, 600, [3, 4], dev)
    for _ in range(n_iter):
        torch.randperm(600, device=dev)  # main.py:502-506
    assert torch.equal(torch.cuda.get_rng_state(0), got_state)
    K = model.total_components
    H0 = torch.rand((K, 200), dtype=torch.float32, device=dev).cpu().numpy()  # the reference's next draw (main.py:687)
    torch.cuda.set_rng_state(got_state, 0)
    val = _adata(n=200, seed=9)
    model.transform(val, n_iter=4)
    W = np.concatenate(model.get_decomposed_matrices()["Ws"], axis=1)
    Ht = orc.transform_loop(np.ascontiguousarray(val.X).T, W, H0, 4, model.eps)
    got = np.concatenate([val.obsm["cov0"].T, val.obsm["cov1"].T, val.obsm["ALPINE_embedding"].T], axis=0)
    assert rel_fro(got, Ht) < 2e-5


@pytest.mark.parametrize("method", ["random", "weighted"])
def test_minibatch_index_stream_equals_the_reference_call_sequence(method, monkeypatch):
    """Un-hooked mini-batch fit: the epoch index vectors equal what the reference's calls yield on freshly seeded
    process-wide generators (device randperm after the init draws; CPU multinomial of the weighted sampler)."""
    from alpine_b200.utils import sampling

    dev = torch.device("cuda:0")
    ad = _adata(n=500, G=300, cats=(3,), nan_fraction=0.0)
    seen = []
    real = sampling.generate_epoch_indices

    def recording(*a, **k):
        idx = real(*a, **k)
        seen.append(idx.cpu().numpy().copy())
        return idx

    monkeypatch.setattr(sampling, "generate_epoch_indices", recording)
    model = ALPINE(n_components=5, n_covariate_components=[3], lam=[1e2], device="cuda:0", random_state=5)
    model.fit(ad, ["cov0"], max_iter=3, batch_size=128, sampling_method=method)
    monkeypatch.setattr(sampling, "generate_epoch_indices", real)
    assert len(seen) == 3
    _reference_draw_sequence(model, 300, 500, [3], dev)
    Y = model.fe.transform(ad.obs)
    joint = sampling.create_joint_labels_from_dummy_matrices([torch.from_numpy(np.ascontiguousarray(y.T)) for y in Y])
    for it in range(3):
        want = real(joint_labels=joint, sampling_method=method, device=dev)  # process-wide generators, as the reference
        np.testing.assert_array_equal(seen[it], want.cpu().numpy())


def test_search_hyperparams_runs_real_fits_on_the_gpu():
    """ComponentOptimizer.search_hyperparams (optimization.py:51-151) driving real fold fits + transforms through the
    device scheduler (4 trials x 2 folds; random-search and k-means stand-ins, as hyperopt / scanpy are absent)."""
    from alpine_b200 import ComponentOptimizer

    ad = _adata(n=900, G=300, cats=(3, 2), nan_fraction=0.0)
    opt = ComponentOptimizer(ad, ["cov0", "cov1"], max_iter=30, device="cuda", random_state=3)
    best = opt.search_hyperparams(n_total_components_range=(10, 24), max_evals=4, n_splits=2)
    assert set(best) == {"n_components", "n_covariate_components", "lam", "alpha_W", "orth_W", "l1_ratio_W", "random_state"}
    hist = opt.get_train_history()
    assert len(hist) >= 1 and np.isfinite(hist["score"]).all()
    assert sum(t > 0 for t in opt.device_busy_s.values()) >= 1
    model = ALPINE(**best, device="cuda")  # the returned dict constructs a model (optimization.py:141-151)
    model.fit(ad, ["cov0", "cov1"], max_iter=5)
    assert len(model.loss_history) == 5
    # same seed, same data => same suggestions and scores (per-model random streams, no cross-thread seeding)
    opt2 = ComponentOptimizer(ad, ["cov0", "cov1"], max_iter=30, device="cuda", random_state=3)
    opt2.search_hyperparams(n_total_components_range=(10, 24), max_evals=4, n_splits=2)
    np.testing.assert_allclose(opt2.get_train_history()["score"].to_numpy(), hist["score"].to_numpy(), rtol=1e-6)


def test_max_iter_zero_and_limits_are_handled_on_the_host():
    ad = _adata(n=200, G=100, cats=(3,), nan_fraction=0.0)
    model = ALPINE(n_components=4, n_covariate_components=[3], lam=[1.0], device="cuda")
    model.fit(ad, ["cov0"], max_iter=0)  # the reference's loop never runs: empty history, scaled initial factors
    assert len(model.loss_history) == 0 and list(model.loss_history.columns)[:2] == ["total loss", "reconstruction loss"]
    with pytest.raises(ValueError, match="at least one component per covariate"):
        ALPINE(n_components=4, n_covariate_components=[0], lam=[1.0], device="cuda").fit(ad, ["cov0"], max_iter=2)


def test_fit_without_covariates_matches_oracle():
    """Plain NMF through the same kernels: no guided blocks, no B, no prediction terms (n_covariate_components=[])."""
    ad = _adata(n=700, G=450, cats=(3,), nan_fraction=0.0)
    kw = dict(n_components=12, n_covariate_components=[], lam=[], orth_W=0.1, alpha_W=0.2, l1_ratio_W=0.5)
    model = ALPINE(device="cuda:0", **kw)
    model.covariate_keys, model.sampling_method, model.verbose = [], "random", False
    X = np.ascontiguousarray(ad.X, dtype=np.float32).T
    model.batch_size, model.max_iter = X.shape[1], 8
    m = model._initialize_matrices(X, [])
    st = orc.State(m.W.cpu().numpy().copy(), m.H.cpu().numpy().copy(), [], [12])
    hp = orc.HyperParams(**kw)
    model._fit(m)
    orc.fit_loop(X, [], st, hp, 8)
    assert rel_fro(m.W.cpu().numpy(), st.W) < 2e-5 and rel_fro(m.H.cpu().numpy(), st.H) < 2e-5
    ref64 = orc.compute_loss(X, [], st, hp, dtype=np.float64)
    lh = model.loss_history
    assert list(lh.columns) == ["total loss", "reconstruction loss"]
    assert abs(lh["reconstruction loss"].iloc[-1] - ref64[1]) / ref64[1] < 1e-4
    out = ALPINE(device="cuda", **kw).fit(ad, [], max_iter=4)
    assert len(out.loss_history) == 4 and out.get_covariate_gene_scores() == {}


def test_fit_on_a_side_stream_equals_the_default_stream():
    """The library launches on the caller's current stream and orders its own initialisation on it: a fit
    under a non-blocking side stream gives the same factors as on the default stream."""
    ad = _adata(n=900, G=500, cats=(3, 4), nan_fraction=0.02)
    keys = ["cov0", "cov1"]
    ref = ALPINE(device="cuda:0", **KW).fit(ad, keys, max_iter=6).get_decomposed_matrices()
    side = torch.cuda.Stream(device="cuda:0")
    side.wait_stream(torch.cuda.current_stream("cuda:0"))
    with torch.cuda.stream(side):
        got = ALPINE(device="cuda:0", **KW).fit(ad, keys, max_iter=6).get_decomposed_matrices()
    side.synchronize()
    for a, b in zip(got["Ws"] + got["Hs"] + got["Bs"], ref["Ws"] + ref["Hs"] + ref["Bs"]):
        np.testing.assert_array_equal(a, b)
