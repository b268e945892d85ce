"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI, against the
oracle (oracle/alpine_oracle.py) and against the golden trajectories the unmodified reference produced.

Tolerances (north_star): per-iteration W/H within 1e-4 Frobenius-relative over the first 10 iterations, final
reconstruction loss within 1e-4 relative, identical top-100 gene rankings.  The measured errors are ~1e-6.
"""
import os

import numpy as np
import pytest
import torch

from oracle import alpine_oracle as orc
from tests.helpers import (CASE_KW, assert_same_top_ranking, full_batch_mu_names, golden_names, hp_of, inputs_of,
                           load_golden, rel_fro)

pytestmark = pytest.mark.gpu

PARITY_TOL = 1e-4      # north_star bar
EXPECTED_TOL = 2e-5    # what fp32 / split-precision arithmetic actually delivers; a regression guard
PRODUCT_TOL = 3e-6     # a single split-precision contraction (tf32 + bf16 corrections) against fp64


def _gpu_utils():
    from tests import gpu_utils

    return gpu_utils


PRODUCT_SHAPES = [
    # (n_cells, n_genes, K)
    (160, 96, 9),
    (203, 132, 16),
    (500, 300, 12),
    (777, 1000, 100),
    (5000, 2000, 25),
    (3001, 2600, 100),
    (1200, 640, 128),
    (40000, 512, 100),
    (1200, 640, 160),    # > 128 components: two launches per contraction (component groups of 80 + 80)
    (900, 500, 256),     # the maximum: 128 + 128
]


@pytest.mark.parametrize("shape", PRODUCT_SHAPES)
def test_contractions_match_fp64(shape):
    """X H^T (main.py:596) and W^T X (main.py:653) as split-precision tcgen05 GEMMs (tf32 hi*hi + bf16 correction
    terms) vs float64 matmul."""
    gu = _gpu_utils()
    n, G, K = shape
    rng = np.random.default_rng(n + G + K)
    X = rng.gamma(0.3, 2.0, size=(n, G)).astype(np.float32)
    W = rng.random((G, K), dtype=np.float32)
    H = rng.random((K, n), dtype=np.float32)
    prob = gu.DeviceProblem(X, [], W, H, [], [K], {})
    xh = prob.solver.xh_product().cpu().numpy().astype(np.float64)   # (K, G)
    wx = prob.solver.wx_product().cpu().numpy().astype(np.float64)   # (K, n)
    Xd, Wd, Hd = X.astype(np.float64), W.astype(np.float64), H.astype(np.float64)
    ref_xh = Hd @ Xd            # (K, n) @ (n, G)
    ref_wx = Wd.T @ Xd.T        # (K, G) @ (G, n)
    assert rel_fro(xh, ref_xh) < PRODUCT_TOL
    assert rel_fro(wx, ref_wx) < PRODUCT_TOL
    # element-wise too: every output is a sum of non-negative terms, so relative error is well defined
    assert np.max(np.abs(xh - ref_xh) / ref_xh) < 2e-5
    assert np.max(np.abs(wx - ref_wx) / ref_wx) < 2e-5


def test_contractions_keep_fp32_range_and_precision_on_wide_data():
    """Operands spanning 25 decades (factor entries from 1e-20 to 1e3, X from 1e-6 to 1e5, plus exact zeros): the
    bf16 images of the correction terms share fp32's exponent range, so nothing over- or underflows and every output
    keeps fp32-level relative accuracy (sums of non-negative terms)."""
    gu = _gpu_utils()
    n, G, K = 1500, 700, 48
    rng = np.random.default_rng(77)
    X = (10.0 ** rng.uniform(-6, 5, size=(n, G))).astype(np.float32)
    X[rng.random((n, G)) < 0.3] = 0.0
    W = (10.0 ** rng.uniform(-20, 3, size=(G, K))).astype(np.float32)
    H = (10.0 ** rng.uniform(-20, 3, size=(K, n))).astype(np.float32)
    prob = gu.DeviceProblem(X, [], W, H, [], [K], {})
    xh = prob.solver.xh_product().cpu().numpy().astype(np.float64)
    wx = prob.solver.wx_product().cpu().numpy().astype(np.float64)
    Xd, Wd, Hd = X.astype(np.float64), W.astype(np.float64), H.astype(np.float64)
    ref_xh, ref_wx = Hd @ Xd, Wd.T @ Xd.T
    assert np.isfinite(xh).all() and np.isfinite(wx).all()
    assert np.max(np.abs(xh - ref_xh) / ref_xh) < 2e-5
    assert np.max(np.abs(wx - ref_wx) / ref_wx) < 2e-5


def test_contraction_is_deterministic():
    gu = _gpu_utils()
    rng = np.random.default_rng(5)
    n, G, K = 4000, 1500, 40
    X = rng.gamma(0.3, 2.0, size=(n, G)).astype(np.float32)
    prob = gu.DeviceProblem(X, [], rng.random((G, K), dtype=np.float32), rng.random((K, n), dtype=np.float32), [], [K], {})
    a = prob.solver.xh_product().clone()
    b = prob.solver.xh_product().clone()
    assert torch.equal(a, b)
    a = prob.solver.wx_product().clone()
    b = prob.solver.wx_product().clone()
    assert torch.equal(a, b)


@pytest.mark.parametrize("exchange_buffer", [False, True])
@pytest.mark.parametrize("name", full_batch_mu_names())
def test_trajectory_matches_reference_golden(name, exchange_buffer):
    """Per-iteration W/H/B against the unmodified reference's trajectory (tests/golden, oracle/gen_golden.py); with
    the W update's numerator summed from the contraction's partial slots (single GPU) and read from the exchange
    buffer (what the multi-GPU engine all-reduces)."""
    gu = _gpu_utils()
    g = load_golden(name)
    kept = [int(i) for i in g["kept_iters"]]
    n_iter = max(kept)
    n_cov = int(g["n_cov"])
    prob = gu.problem_from_golden(name, g, exchange_buffer=exchange_buffer)
    tol = PARITY_TOL
    seen = {}

    def check(it):
        if it in kept and it <= 10:
            W, H, Bs = prob.host()
            seen[it] = (rel_fro(W, g[f"W_it{it}"]), rel_fro(H, g[f"H_it{it}"]))
            assert seen[it][0] < tol, (name, it, "W", seen[it])
            assert seen[it][1] < tol, (name, it, "H", seen[it])
            assert max(seen[it]) < EXPECTED_TOL, (name, it, seen[it])
            for i in range(n_cov):
                assert rel_fro(Bs[i], g[f"B{i}_it{it}"]) < tol, (name, it, f"B{i}")

    xn, rows = prob.run(n_iter, on_iter=check)
    W, H, Bs = prob.host()
    it = kept[-1]
    long_tol = PARITY_TOL if n_iter <= 10 else 2e-4
    assert rel_fro(W, g[f"W_it{it}"]) < long_tol
    assert rel_fro(H, g[f"H_it{it}"]) < long_tol
    # final reconstruction loss by the trace identity vs the fp64 re-evaluation of the reference's final factors
    recon = xn - 2.0 * rows[-1, 0] + rows[-1, 1]
    assert abs(recon - float(g["final_recon_fp64"])) / float(g["final_recon_fp64"]) < PARITY_TOL
    # prediction losses against the reference's own loss history (fp32 on CPU)
    ref_hist = g["loss_history_ref_fp32"]
    for i in range(n_cov):
        np.testing.assert_allclose(rows[-1, 2 + i], ref_hist[n_iter - 1][2 + i], rtol=1e-3, atol=1e-6 * prob.solver.n)
    # every iteration's loss, against the reference's fp32 history (its own error is ~6e-4)
    recon_hist = xn - 2.0 * rows[:, 0] + rows[:, 1]
    np.testing.assert_allclose(recon_hist, ref_hist[:, 1], rtol=2e-3)


@pytest.mark.parametrize("sparse", [False, True])
def test_als_trajectory_matches_reference_golden(sparse):
    """use_als=True (block Gauss-Seidel sweep, main.py:523-588) against the unmodified reference's trajectory."""
    gu = _gpu_utils()
    name = "als_reg"
    g = load_golden(name)
    kept = [int(i) for i in g["kept_iters"]]
    n_cov = int(g["n_cov"])
    prob = gu.problem_from_golden(name, g, sparse=sparse)

    def check(it):
        if it in kept:
            W, H, Bs = prob.host()
            assert rel_fro(W, g[f"W_it{it}"]) < EXPECTED_TOL, (it, "W")
            assert rel_fro(H, g[f"H_it{it}"]) < EXPECTED_TOL, (it, "H")
            for i in range(n_cov):
                assert rel_fro(Bs[i], g[f"B{i}_it{it}"]) < PARITY_TOL, (it, f"B{i}")

    n_iter = max(kept)
    xn, rows = prob.run(n_iter, on_iter=check, use_als=True)
    recon = xn - 2.0 * rows[-1, 0] + rows[-1, 1]
    assert abs(recon - float(g["final_recon_fp64"])) / float(g["final_recon_fp64"]) < PARITY_TOL
    ref_hist = g["loss_history_ref_fp32"]
    np.testing.assert_allclose(xn - 2.0 * rows[:, 0] + rows[:, 1], ref_hist[:, 1], rtol=2e-3)
    for i in range(n_cov):
        np.testing.assert_allclose(rows[:, 2 + i], ref_hist[:, 2 + i], rtol=1e-3, atol=1e-6 * prob.solver.n)


def test_als_matches_oracle_at_wider_blocks():
    """ALS with component blocks wider than one MMA N-group (k_b = 40, 24, 36) and a ragged cell count."""
    gu = _gpu_utils()
    from alpine_b200.utils.synth import labels_to_dummies, make_counts, make_labels

    n, G, blocks = 1531, 900, [40, 24, 36]
    kw = dict(n_components=36, n_covariate_components=[40, 24], lam=[1e2, 3e2], orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5)
    X = make_counts(n, G, seed=4, rank=10)
    Ycg, _ = labels_to_dummies(make_labels(n, [5, 3], seed=4, nan_fraction=0.02))
    Ys = [np.ascontiguousarray(y.T) for y in Ycg]
    rng = np.random.default_rng(1)
    K = sum(blocks)
    W0 = np.maximum(rng.random((G, K), dtype=np.float32), 1e-6)
    H0 = np.maximum(rng.random((K, n), dtype=np.float32), 1e-6)
    B0 = [np.maximum(rng.random((c, k), dtype=np.float32), 1e-6) for c, k in zip([5, 3], blocks)]
    hp = orc.HyperParams(**kw)
    st = orc.State(W0.copy(), H0.copy(), [b.copy() for b in B0], blocks)
    prob = gu.DeviceProblem(X, Ys, W0, H0, B0, blocks, kw)

    def check(it):
        orc.als_step(X.T, Ys, st, hp)
        W, H, Bs = prob.host()
        assert max(rel_fro(W, st.W), rel_fro(H, st.H)) < EXPECTED_TOL, it
        assert max(rel_fro(a, b) for a, b in zip(Bs, st.Bs)) < PARITY_TOL, it

    xn, rows = prob.run(6, on_iter=check, use_als=True)
    ref = orc.compute_loss(X.T, Ys, st, hp, dtype=np.float64)
    assert abs((xn - 2.0 * rows[-1, 0] + rows[-1, 1]) - ref[1]) / ref[1] < PARITY_TOL


def test_trajectory_matches_oracle_cfg1_shapes():
    """BASELINE config[0] shapes (2,000 genes x 5,000 cells, 20+[5] components) for 10 iterations vs the oracle."""
    gu = _gpu_utils()
    from alpine_b200.utils.synth import labels_to_dummies, make_counts, make_labels

    n, G = 5000, 2000
    kw = dict(n_components=20, n_covariate_components=[5], lam=[1e3], orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5)
    X = make_counts(n, G, seed=0, rank=8)
    Ycg, _ = labels_to_dummies(make_labels(n, [3], seed=0))
    Ys = [np.ascontiguousarray(y.T) for y in Ycg]
    rng = np.random.default_rng(42)
    blocks = [5, 20]
    K = sum(blocks)
    W0 = np.maximum(rng.random((G, K), dtype=np.float32), 1e-6)
    H0 = np.maximum(rng.random((K, n), dtype=np.float32), 1e-6)
    B0 = [np.maximum(rng.random((3, 5), dtype=np.float32), 1e-6)]
    hp = orc.HyperParams(**kw)
    st = orc.State(W0.copy(), H0.copy(), [b.copy() for b in B0], blocks)
    Xg = X.T  # genes x cells view, as the reference holds it
    prob = gu.DeviceProblem(X, Ys, W0, H0, B0, blocks, kw)
    worst = [0.0]

    def check(it):
        orc.mu_step(Xg, Ys, st, hp)
        W, H, Bs = prob.host()
        e = max(rel_fro(W, st.W), rel_fro(H, st.H), rel_fro(Bs[0], st.Bs[0]))
        worst[0] = max(worst[0], e)
        assert e < PARITY_TOL, (it, e)

    xn, rows = prob.run(10, on_iter=check)
    assert worst[0] < EXPECTED_TOL
    ref = orc.compute_loss(Xg, Ys, st, hp, dtype=np.float64)
    recon = xn - 2.0 * rows[-1, 0] + rows[-1, 1]
    assert abs(recon - ref[1]) / ref[1] < PARITY_TOL
    assert abs(rows[-1, 2] - ref[2]) <= 1e-3 * abs(ref[2]) + 1e-6 * n


def test_long_run_top100_rankings_match_reference():
    """200 iterations: identical top-100 genes per component (north_star ranking criterion)."""
    gu = _gpu_utils()
    g = load_golden("kl_long200")
    prob = gu.problem_from_golden("kl_long200", g)
    prob.run(200)
    W, H, _ = prob.host()
    Wref = g["W_it200"]
    assert rel_fro(W, Wref) < 2e-4
    for k in range(W.shape[1]):
        assert_same_top_ranking(W[:, k], Wref[:, k], what=("W column", k))


@pytest.mark.parametrize("name", full_batch_mu_names())
def test_scale_and_transform_match_reference_golden(name):
    gu = _gpu_utils()
    g = load_golden(name)
    n_cov = int(g["n_cov"])
    it = int(g["kept_iters"][-1])
    kw = dict(CASE_KW[name])
    Ys = [np.ascontiguousarray(g[f"Y{i}_cells_by_cat"].T) for i in range(n_cov)]
    prob = gu.DeviceProblem(g["X_cells_by_genes"], Ys, g[f"W_it{it}"], g[f"H_it{it}"],
                            [g[f"B{i}_it{it}"] for i in range(n_cov)], [int(b) for b in g["blocks"]], kw)
    prob.solver.scale()
    W, H, Bs = prob.host()
    assert rel_fro(W, g["W_scaled"]) < 1e-6
    assert rel_fro(H, g["H_scaled"]) < 1e-6
    for i in range(n_cov):
        assert rel_fro(Bs[i], g[f"B{i}_scaled"]) < 1e-6
    # H-only transform (main.py:705-709) with the scaled W and the fixture's H0
    prob2 = gu.DeviceProblem(g["X_cells_by_genes"], [], g["W_scaled"], g["Ht0"], [], [W.shape[1]], {})
    prob2.solver.transform(5)
    assert rel_fro(prob2.H.cpu().numpy(), g["Ht_5"]) < EXPECTED_TOL


def test_argument_errors_are_reported():
    from alpine_b200 import _native

    with pytest.raises(_native.AlpineNativeError):
        _native.Solver("cuda:0", 10, 10, [257], [])  # K > 256
    s = _native.Solver("cuda:0", 64, 64, [4], [])
    with pytest.raises(_native.AlpineNativeError):
        s.fit_begin(3)  # nothing bound


TINY_SHAPES = [
    # n_cells, n_genes, blocks (covariate blocks first), categories
    (5, 3, [1, 1], [2]),
    (1, 7, [2, 1], [2]),
    (33, 257, [3, 14], [4]),
    (64, 1, [2, 2], [3]),
    (300, 40, [64, 64], [5]),       # K = 128: both TMEM accumulators full width, widest guided block
    (300, 40, [70, 130], [5]),      # K = 200: two component groups per contraction (no block-wise sweep above 128)
]


@pytest.mark.parametrize("shape", TINY_SHAPES)
@pytest.mark.parametrize("mode", ["mu", "als", "csr"])
def test_degenerate_and_maximal_shapes_match_oracle(shape, mode):
    """Shapes far below one tile (TMA boxes mostly out of bounds), a single cell / gene, and the K = 128 limit."""
    gu = _gpu_utils()
    n, G, blocks, cats = shape
    if mode == "als" and sum(blocks) > 128:
        pytest.skip("use_als supports at most 128 components in total")
    rng = np.random.default_rng(n * 1000 + G)
    X = rng.gamma(0.5, 2.0, size=(n, G)).astype(np.float32)
    X[rng.random((n, G)) < 0.3] = 0.0
    K = sum(blocks)
    lab = rng.integers(0, cats[0], n)
    Ys = [np.ascontiguousarray(np.eye(cats[0], dtype=np.float32)[lab].T)]
    W0 = np.maximum(rng.random((G, K), dtype=np.float32), 1e-6)
    H0 = np.maximum(rng.random((K, n), dtype=np.float32), 1e-6)
    B0 = [np.maximum(rng.random((cats[0], blocks[0]), dtype=np.float32), 1e-6)]
    kw = dict(n_components=blocks[-1], n_covariate_components=blocks[:-1], lam=[10.0], orth_W=0.1, alpha_W=0.2,
              l1_ratio_W=0.5)
    hp = orc.HyperParams(**kw)
    st = orc.State(W0.copy(), H0.copy(), [b.copy() for b in B0], blocks)
    prob = gu.DeviceProblem(X, Ys, W0, H0, B0, blocks, kw, sparse=(mode == "csr"))
    step = orc.als_step if mode == "als" else orc.mu_step

    def check(it):
        step(X.T, Ys, st, hp)
        W, H, Bs = prob.host()
        assert max(rel_fro(W, st.W), rel_fro(H, st.H), rel_fro(Bs[0], st.Bs[0])) < PARITY_TOL, (shape, mode, it)

    xn, rows = prob.run(4, on_iter=check, use_als=(mode == "als"))
    ref = orc.compute_loss(X.T, Ys, st, hp, dtype=np.float64)
    recon = xn - 2.0 * rows[-1, 0] + rows[-1, 1]
    # the trace identity subtracts quantities of size ||X||^2 held in fp32 factors: its absolute error is ~1e-7 ||X||^2,
    # which only shows when the fit is near-exact (a single cell is a rank-1 problem: recon / ||X||^2 = 7e-5)
    assert abs(recon - ref[1]) <= 1e-4 * ref[1] + 1e-6 * xn


def test_peer_exchange_fallback_keeps_the_all_reduce_path_working():
    """A rank that exported its exchange block but then falls back (another rank could not map the peers) must run
    the ordinary path on the block's storage (W^T lives inside it) with unchanged results."""
    import ctypes

    gu = _gpu_utils()
    from alpine_b200 import _native

    name = "kl_reg_nan"
    g = load_golden(name)
    n_cov = int(g["n_cov"])
    Ys = [np.ascontiguousarray(g[f"Y{i}_cells_by_cat"].T) for i in range(n_cov)]
    kw = dict(CASE_KW[name])
    dev = torch.device("cuda:0")
    n, G = g["X_cells_by_genes"].shape
    X = gu.to_dev_padded(g["X_cells_by_genes"], dev)
    W = torch.from_numpy(g["W0"].copy()).to(dev)
    H = gu.to_dev_padded(g["H0"], dev)
    Yd = [torch.from_numpy(y).to(dev) for y in Ys]
    Bd = [torch.from_numpy(g[f"B0_{i}"].copy()).to(dev) for i in range(n_cov)]
    s = _native.Solver(dev, G, n, [int(b) for b in g["blocks"]], [y.shape[0] for y in Ys])
    handle = (ctypes.c_ubyte * 64)()
    assert s.lib.alpine_peer_export(s._ctx, ctypes.cast(handle, ctypes.c_void_p)) == 0
    assert s.lib.alpine_peer_disable(s._ctx) == 0
    s.bind_dense(X)
    s.bind_labels(Yd)
    s.bind_factors(W, H, Bd)
    s.set_hparams(kw["lam"], kw["alpha_W"], kw["l1_ratio_W"], kw["orth_W"], 1e-6)
    with pytest.raises(_native.AlpineNativeError):
        s.fit_begin(3)
        s.mu_partials()
        s.mu_apply_peer(0)  # no peers were imported
    s.fit_begin(5)
    for it in range(5):
        s.mu_partials()
        s.mu_apply(it)
    s.losses(5)
    assert rel_fro(W.cpu().numpy(), g["W_it5"]) < EXPECTED_TOL
    assert rel_fro(H.cpu().numpy(), g["H_it5"]) < EXPECTED_TOL
    s.close()
