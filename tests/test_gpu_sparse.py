"""GPU tests of the CSR (tile-list) variant of the contractions: the same kernels with a sparse X producer
(csrc/csr_tiles.cuh, mu_gemm_sm100.cuh SRC_TILES).  The reference has no sparse path (main.py:395-396 rejects it),
so parity is against the dense path / the oracle on the densified matrix: identical mathematics."""
import numpy as np
import pandas as pd
import pytest
import torch

from oracle import alpine_oracle as orc
from tests.helpers import full_batch_mu_names, load_golden, rel_fro

pytestmark = pytest.mark.gpu


def _gu():
    from tests import gpu_utils

    return gpu_utils


def _sparse_counts(n, G, density, seed, integer=True):
    rng = np.random.default_rng(seed)
    mask = rng.random((n, G)) < density
    vals = (1 + rng.poisson(2.0, size=(n, G))).astype(np.float32) if integer else rng.gamma(0.3, 2.0, size=(n, G)).astype(np.float32)
    return np.where(mask, vals, 0).astype(np.float32)


SHAPES = [
    # n_cells, n_genes, K, density, integer values
    (160, 96, 9, 0.3, True),
    (777, 1000, 100, 0.05, True),
    (3001, 2600, 100, 0.05, False),
    (5000, 2000, 25, 0.02, False),
    (1200, 640, 128, 0.9, False),     # denser than the 512 entries per k-block held in registers
    (40000, 512, 100, 0.05, True),
    (300, 40000, 40, 0.01, True),
]


@pytest.mark.parametrize("shape", SHAPES)
def test_csr_contractions_match_dense_and_fp64(shape):
    gu = _gu()
    n, G, K, density, integer = shape
    X = _sparse_counts(n, G, density, seed=n + G, integer=integer)
    rng = np.random.default_rng(K)
    W = rng.random((G, K), dtype=np.float32)
    H = rng.random((K, n), dtype=np.float32)
    dense = gu.DeviceProblem(X, [], W, H, [], [K], {})
    sparse = gu.DeviceProblem(X, [], W, H, [], [K], {}, sparse=True)
    xh_s, wx_s = sparse.solver.xh_product(), sparse.solver.wx_product()
    Xd, Wd, Hd = X.astype(np.float64), W.astype(np.float64), H.astype(np.float64)
    assert rel_fro(xh_s.cpu().numpy(), Hd @ Xd) < 3e-6
    assert rel_fro(wx_s.cpu().numpy(), Wd.T @ Xd.T) < 3e-6
    if not integer:
        # same tiles, same work split, same MMAs: the sparse producer must reproduce the dense path bit for bit
        # (with integer counts the sparse context knows X is tf32-exact at bind time and skips the lo products;
        # the dense context only learns that in fit_begin)
        assert torch.equal(xh_s, dense.solver.xh_product())
        assert torch.equal(wx_s, dense.solver.wx_product())


def test_csr_empty_rows_and_empty_matrix_blocks():
    gu = _gu()
    n, G, K = 900, 700, 12
    X = _sparse_counts(n, G, 0.05, seed=3)
    X[100:400] = 0          # cells without any count (whole super-tile rows of zeros)
    X[:, 256:512] = 0       # a gene super-tile without nonzeros
    rng = np.random.default_rng(0)
    W, H = rng.random((G, K), dtype=np.float32), rng.random((K, n), dtype=np.float32)
    p = gu.DeviceProblem(X, [], W, H, [], [K], {}, sparse=True)
    xh, wx = p.solver.xh_product().cpu().numpy(), p.solver.wx_product().cpu().numpy()
    assert rel_fro(xh, H.astype(np.float64) @ X.astype(np.float64)) < 3e-6
    assert rel_fro(wx, W.astype(np.float64).T @ X.astype(np.float64).T) < 3e-6
    assert np.all(wx[:, 100:400] == 0) and np.all(xh[:, 256:512] == 0)


def test_csr_bad_input_is_reported():
    from alpine_b200 import _native

    s = _native.Solver("cuda:0", 50, 4, [3], [])
    dev = torch.device("cuda:0")
    indptr = torch.tensor([0, 1, 2, 2, 3], dtype=torch.int64, device=dev)
    vals = torch.ones(3, dtype=torch.float32, device=dev)
    with pytest.raises(_native.AlpineNativeError, match="column index"):
        s.bind_csr(indptr, torch.tensor([0, 50, 3], dtype=torch.int32, device=dev), vals)
    with pytest.raises(_native.AlpineNativeError, match="non-negative"):
        s.bind_csr(indptr, torch.tensor([0, 5, 3], dtype=torch.int32, device=dev), -vals)
    with pytest.raises(_native.AlpineNativeError, match="nnz"):
        s.bind_csr(torch.tensor([0, 1, 2, 2, 2], dtype=torch.int64, device=dev),
                   torch.tensor([0, 5, 3], dtype=torch.int32, device=dev), vals)


@pytest.mark.parametrize("name", full_batch_mu_names()[:3])
def test_csr_trajectory_matches_reference_golden(name):
    """The golden trajectories of the unmodified reference, with X handed over as CSR."""
    gu = _gu()
    g = load_golden(name)
    kept = [int(i) for i in g["kept_iters"] if int(i) <= 10]
    n_cov = int(g["n_cov"])
    prob = gu.problem_from_golden(name, g, sparse=True)

    def check(it):
        if it in kept:
            W, H, Bs = prob.host()
            assert rel_fro(W, g[f"W_it{it}"]) < 2e-5, (name, it, "W")
            assert rel_fro(H, g[f"H_it{it}"]) < 2e-5, (name, it, "H")
            for i in range(n_cov):
                assert rel_fro(Bs[i], g[f"B{i}_it{it}"]) < 1e-4, (name, it, f"B{i}")

    n_iter = max(kept)
    xn, rows = prob.run(n_iter, on_iter=check)
    if n_iter == int(g["kept_iters"][-1]):
        recon = xn - 2.0 * rows[-1, 0] + rows[-1, 1]
        assert abs(recon - float(g["final_recon_fp64"])) / float(g["final_recon_fp64"]) < 1e-4


def test_count_matrix_fast_path_dense_and_sparse_agree_with_oracle():
    """Integer counts are tf32-exact: both paths drop the lo half of the X split (2 MMAs per k-step)."""
    gu = _gu()
    from alpine_b200.utils.synth import labels_to_dummies, make_labels

    n, G, blocks = 2000, 1500, [4, 16]
    X = _sparse_counts(n, G, 0.08, seed=11)
    Ycg, _ = labels_to_dummies(make_labels(n, [3], seed=1))
    Ys = [np.ascontiguousarray(y.T) for y in Ycg]
    rng = np.random.default_rng(42)
    K = sum(blocks)
    W0 = np.maximum(rng.random((G, K), dtype=np.float32), 1e-6)
    H0 = np.maximum(rng.random((K, n), dtype=np.float32), 1e-6)
    B0 = [np.maximum(rng.random((3, 4), dtype=np.float32), 1e-6)]
    kw = dict(n_components=16, n_covariate_components=[4], lam=[1e2], orth_W=0.1, alpha_W=0.2, l1_ratio_W=0.5)
    hp = orc.HyperParams(**kw)
    st = orc.State(W0.copy(), H0.copy(), [b.copy() for b in B0], blocks)
    for _ in range(5):
        orc.mu_step(X.T, Ys, st, hp)
    outs = []
    for sparse in (False, True):
        prob = gu.DeviceProblem(X, Ys, W0, H0, B0, blocks, kw, sparse=sparse)
        xn, rows = prob.run(5)
        W, H, Bs = prob.host()
        assert rel_fro(W, st.W) < 2e-5 and rel_fro(H, st.H) < 2e-5 and rel_fro(Bs[0], st.Bs[0]) < 2e-5
        assert abs(xn - float((X.astype(np.float64) ** 2).sum())) < 1e-9 * xn
        outs.append((W, H))
    # after fit_begin the dense context has seen that X is exact too: identical arithmetic from then on
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_public_fit_with_scipy_csr_matches_dense_fit():
    import scipy.sparse as sp

    from alpine_b200 import ALPINE
    from alpine_b200.utils.anndata_compat import AnnData
    from alpine_b200.utils.synth import make_labels

    n, G = 700, 500
    X = _sparse_counts(n, G, 0.1, seed=5)
    labels = make_labels(n, [3], seed=2)
    obs = pd.DataFrame({"cov0": pd.Series(labels[0], dtype=object)})
    obs.index = [f"c{i}" for i in range(n)]
    var = pd.DataFrame(index=[f"g{i}" for i in range(G)])
    kw = dict(n_components=6, n_covariate_components=[3], lam=[1e2], alpha_W=0.1, device="cuda")
    a_dense, a_sparse = AnnData(X.copy(), obs=obs.copy(), var=var.copy()), AnnData(sp.csr_matrix(X), obs=obs.copy(), var=var.copy())
    m_dense = ALPINE(**kw).fit(a_dense, ["cov0"], max_iter=8)
    m_sparse = ALPINE(**kw).fit(a_sparse, ["cov0"], max_iter=8)
    np.testing.assert_allclose(a_sparse.obsm["ALPINE_embedding"], a_dense.obsm["ALPINE_embedding"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(a_sparse.varm["cov0"], a_dense.varm["cov0"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(m_sparse.loss_history.to_numpy(), m_dense.loss_history.to_numpy(), rtol=1e-6)
    val = AnnData(sp.csr_matrix(X[:200]), obs=obs.iloc[:200].copy(), var=var.copy())
    torch.manual_seed(3)
    m_sparse.transform(val, n_iter=4)
    assert val.obsm["ALPINE_embedding"].shape == (200, 6) and np.isfinite(val.obsm["ALPINE_embedding"]).all()
    assert np.isfinite(m_sparse.compute_loss(val))
