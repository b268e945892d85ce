"""GPU parity at the shapes BASELINE.json names (run on the B200 box with `-m gpu`).

The small fixtures under tests/golden/ pin the arithmetic; these tests pin it at sizes where every axis spans many
256-row super-tiles, stream-K pieces and partial-sum slots:

* K = 100 (the headline component count, 5 + 5 guided + 90 unguided) at 3,001 cells x 2,600 genes for 10 iterations
  against the NumPy oracle;
* cfg 2 (5,000 genes x 50,000 cells, 30 + [5, 5]) for 10 iterations and cfg 3 (20,000 genes x 100,000 cells,
  K = 100) for 3 iterations against ``oracle/torch_port.py`` run on ``device="cuda"`` with TF32 switched off --
  the reference's own GPU arithmetic (torch operators of main.py:589-663 on cuBLAS fp32), pinned on the
  reference-generated fixtures by tests/test_oracle_golden.py.  W / H within 1e-4 Frobenius-relative at every
  iteration, final reconstruction loss within 1e-4 of the fp64 re-evaluation ||X - W H||^2 of the port's factors
  (the reference's fp32 torch.norm is itself only good to ~6e-4);
* top-100 rankings of W columns and of ``get_covariate_gene_scores`` (main.py:246-273) on a 2,000-gene
  count-matrix fixture produced by the unmodified reference.
"""
import numpy as np
import pandas as pd
import pytest
import torch

from oracle import alpine_oracle as orc
from oracle import torch_port as tp
from tests.helpers import CASE_KW, assert_same_top_ranking, load_golden, rel_fro

pytestmark = pytest.mark.gpu

PARITY_TOL = 1e-4
EXPECTED_TOL = 2e-5


def _gpu_utils():
    from tests import gpu_utils

    return gpu_utils


def test_k100_simultaneous_update_matches_oracle_over_many_tiles():
    """The headline configuration (K = 100, non-ALS, guided blocks with NaN labels, all regularisers) at a shape with
    12 x 11 super-tiles, ragged in both axes."""
    gu = _gpu_utils()
    from alpine_b200.utils.synth import labels_to_dummies, make_counts, make_labels

    n, G, blocks, cats = 3001, 2600, [5, 5, 90], [3, 4]
    kw = dict(n_components=90, n_covariate_components=[5, 5], lam=[1e3, 1e3], orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5)
    X = make_counts(n, G, seed=11, rank=12)
    Ycg, _ = labels_to_dummies(make_labels(n, cats, seed=11, nan_fraction=0.03))
    Ys = [np.ascontiguousarray(y.T) for y in Ycg]
    rng = np.random.default_rng(42)
    K = sum(blocks)
    W0 = np.maximum(rng.random((G, K), dtype=np.float32), 1e-6)
    H0 = np.maximum(rng.random((K, n), dtype=np.float32), 1e-6)
    B0 = [np.maximum(rng.random((c, k), dtype=np.float32), 1e-6) for c, k in zip(cats, blocks)]
    hp = orc.HyperParams(**kw)
    st = orc.State(W0.copy(), H0.copy(), [b.copy() for b in B0], blocks)
    prob = gu.DeviceProblem(X, Ys, W0, H0, B0, blocks, kw)
    worst = [0.0]

    def check(it):
        orc.mu_step(X.T, Ys, st, hp)
        W, H, Bs = prob.host()
        e = max(rel_fro(W, st.W), rel_fro(H, st.H), max(rel_fro(a, b) for a, b in zip(Bs, st.Bs)))
        worst[0] = max(worst[0], e)
        assert e < PARITY_TOL, (it, e)

    xn, rows = prob.run(10, on_iter=check)
    assert worst[0] < EXPECTED_TOL
    ref = orc.compute_loss(X.T, Ys, st, hp, dtype=np.float64)
    recon = xn - 2.0 * rows[-1, 0] + rows[-1, 1]
    assert abs(recon - ref[1]) / ref[1] < PARITY_TOL
    for i in range(2):
        assert abs(rows[-1, 2 + i] - ref[2 + i]) <= 1e-3 * abs(ref[2 + i]) + 1e-6 * n


@pytest.mark.parametrize("exchange_buffer", [False, True])
def test_two_hundred_components_match_oracle(exchange_buffer):
    """K = 200 (> 128: every contraction runs as two launches over component groups of 112 + 88, the update kernels
    read both groups' slots) against the NumPy oracle; the reference has no limit on K (main.py:47-80)."""
    gu = _gpu_utils()
    from alpine_b200.utils.synth import labels_to_dummies, make_counts, make_labels

    n, G, blocks, cats = 2050, 1300, [6, 4, 190], [3, 4]
    kw = dict(n_components=190, n_covariate_components=[6, 4], lam=[1e3, 5e2], orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5)
    X = make_counts(n, G, seed=21, rank=12)
    Ycg, _ = labels_to_dummies(make_labels(n, cats, seed=21, nan_fraction=0.02))
    Ys = [np.ascontiguousarray(y.T) for y in Ycg]
    rng = np.random.default_rng(7)
    K = sum(blocks)
    W0 = np.maximum(rng.random((G, K), dtype=np.float32), 1e-6)
    H0 = np.maximum(rng.random((K, n), dtype=np.float32), 1e-6)
    B0 = [np.maximum(rng.random((c, k), dtype=np.float32), 1e-6) for c, k in zip(cats, blocks)]
    hp = orc.HyperParams(**kw)
    st = orc.State(W0.copy(), H0.copy(), [b.copy() for b in B0], blocks)
    prob = gu.DeviceProblem(X, Ys, W0, H0, B0, blocks, kw, exchange_buffer=exchange_buffer)

    def check(it):
        orc.mu_step(X.T, Ys, st, hp)
        W, H, Bs = prob.host()
        e = max(rel_fro(W, st.W), rel_fro(H, st.H), max(rel_fro(a, b) for a, b in zip(Bs, st.Bs)))
        assert e < EXPECTED_TOL, (it, e)

    xn, rows = prob.run(5, on_iter=check)
    ref = orc.compute_loss(X.T, Ys, st, hp, dtype=np.float64)
    assert abs((xn - 2.0 * rows[-1, 0] + rows[-1, 1]) - ref[1]) / ref[1] < PARITY_TOL
    # the H-only transform with the same component groups
    prob2 = gu.DeviceProblem(X, [], st.W, H0, [], [K], {})
    prob2.solver.transform(3)
    Ht = orc.transform_loop(X.T, st.W, H0, 3, 1e-6)
    assert rel_fro(prob2.H.cpu().numpy(), Ht) < EXPECTED_TOL


def _recon_fp64(Xcm: torch.Tensor, W: torch.Tensor, H: torch.Tensor, chunk: int = 8192) -> float:
    """||X - W H||_F^2 in fp64, over chunks of cells (X is cells-major: Xcm[j][g])."""
    Wd = W.double()
    total = 0.0
    for j0 in range(0, Xcm.shape[0], chunk):
        j1 = min(Xcm.shape[0], j0 + chunk)
        R = Xcm[j0:j1].double() - (Wd @ H[:, j0:j1].double()).T
        total += float((R * R).sum())
    return total


def _device_inputs(n, G, blocks, cats, dev, seed):
    """Low-rank + noise X (cells-major), labels with 2 % missing, random factors -- all generated on the device."""
    from alpine_b200 import _native

    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    K = sum(blocks)
    Wg = torch.rand((16, G), device=dev, generator=g).pow_(3.0)
    X = _native.padded_rows(n, G, dev)
    step = max(1, (1 << 27) // G)
    for r0 in range(0, n, step):
        r1 = min(n, r0 + step)
        blk = torch.rand((r1 - r0, 16), device=dev, generator=g).pow_(2.0) @ Wg
        blk.add_(torch.rand((r1 - r0, G), device=dev, generator=g).pow_(4.0), alpha=0.5)
        X[r0:r1] = blk
    Ys = []
    for c in cats:
        codes = torch.randint(0, c, (n,), device=dev, generator=g)
        y = torch.nn.functional.one_hot(codes, c).T.contiguous().float()
        y[:, torch.rand((n,), device=dev, generator=g) < 0.02] = 0.0  # missing labels: all-zero column (encoder.py:32-37)
        Ys.append(y)
    W = torch.rand((G, K), device=dev, generator=g).clamp_(min=1e-6)
    H = _native.padded_rows(K, n, dev)
    H.copy_(torch.rand((K, n), device=dev, generator=g).clamp_(min=1e-6))
    Bs = [torch.rand((c, k), device=dev, generator=g).clamp_(min=1e-6).contiguous() for c, k in zip(cats, blocks)]
    return X, Ys, W, H, Bs


BIG_CASES = {
    # BASELINE.json configs[1] and configs[2]
    "cfg2": dict(n=50_000, G=5_000, blocks=[5, 5, 30], cats=[3, 4], n_iter=10,
                 kw=dict(n_components=30, n_covariate_components=[5, 5], lam=[1e3, 1e3])),
    "cfg3": dict(n=100_000, G=20_000, blocks=[5, 5, 90], cats=[3, 4], n_iter=3,
                 kw=dict(n_components=90, n_covariate_components=[5, 5], lam=[1e3, 1e3], orth_W=0.2, alpha_W=0.5,
                         l1_ratio_W=0.5)),
}


@pytest.mark.parametrize("case", list(BIG_CASES))
def test_named_config_trajectory_matches_torch_cuda_port(case):
    from alpine_b200 import _native

    spec = BIG_CASES[case]
    dev = torch.device("cuda:0")
    free, _ = torch.cuda.mem_get_info(dev)
    need = 6.5 * 4.0 * spec["n"] * spec["G"]  # X + the port's G x n temporaries (2X, (2W)H, WH, X - WH) + slack
    if free < need:
        pytest.skip(f"{case}: needs {need / 2**30:.0f} GiB of free HBM")
    n, G, blocks, cats, kw = spec["n"], spec["G"], spec["blocks"], spec["cats"], spec["kw"]
    X, Ys, W, H, Bs = _device_inputs(n, G, blocks, cats, dev, seed=5)
    # the port's own copies of the factors (ours are updated in place); X is shared: genes x cells view, the strides the
    # reference's tensor has (main.py:104, 445)
    Wp, Hp, Bp = W.clone(), H.clone().contiguous(), [b.clone() for b in Bs]
    Xg = X.T
    hp = orc.HyperParams(**kw)
    old_tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False  # the reference's fp32 GEMMs
    s = _native.Solver(dev, G, n, blocks, cats)
    try:
        s.bind_dense(X)
        s.bind_labels(Ys)
        s.bind_factors(W, H, Bs)
        s.set_hparams(kw["lam"], kw.get("alpha_W", 0.0), kw.get("l1_ratio_W", 0.0), kw.get("orth_W", 0.0), 1e-6)
        n_iter = spec["n_iter"]
        s.fit_begin(n_iter)
        worst = 0.0
        for it in range(n_iter):
            s.mu_partials()
            s.mu_apply(it)
            s.sync_w()
            tp.mu_step(Xg, Ys, Wp, Hp, Bp, blocks, hp)
            eW = float(torch.linalg.norm((W - Wp).double()) / torch.linalg.norm(Wp.double()))
            eH = float(torch.linalg.norm((H - Hp).double()) / torch.linalg.norm(Hp.double()))
            eB = max(float(torch.linalg.norm((a - b).double()) / torch.linalg.norm(b.double())) for a, b in zip(Bs, Bp))
            worst = max(worst, eW, eH, eB)
            assert max(eW, eH, eB) < PARITY_TOL, (case, it + 1, eW, eH, eB)
        xn, rows = s.losses(n_iter)
    finally:
        s.close()
        torch.backends.cuda.matmul.allow_tf32 = old_tf32
    assert worst < EXPECTED_TOL, (case, worst)
    recon = xn - 2.0 * rows[-1, 0] + rows[-1, 1]
    ref = _recon_fp64(X, Wp, Hp)
    assert abs(recon - ref) / ref < PARITY_TOL, (case, recon, ref)
    # prediction terms against the port's fp32 evaluation of main.py:727-748
    sls = tp.block_slices(blocks)
    for i, Y in enumerate(Ys):
        y_hat = torch.clamp(Bp[i] @ Hp[sls[i]], min=hp.eps)
        pred = float(torch.sum(Y * torch.log(torch.clamp(Y / y_hat, min=hp.eps)) - Y + y_hat).double())
        assert abs(rows[-1, 2 + i] - pred) <= 1e-3 * abs(pred) + 1e-6 * n, (case, i)


def test_top100_rankings_of_W_and_gene_scores_match_reference_golden():
    """60 iterations on the 2,000-gene count fixture (tf32-exact X: the 2-MMA kernel variant), scaling, then
    ``get_covariate_gene_scores`` on the GPU output: identical top-100 genes per W column and per category."""
    gu = _gpu_utils()
    from alpine_b200 import ALPINE
    from alpine_b200.utils.encoder import FeatureEncoders

    name = "kl_scores2k"
    g = load_golden(name)
    n_cov = int(g["n_cov"])
    it = int(g["kept_iters"][-1])
    prob = gu.problem_from_golden(name, g)
    prob.run(it)
    W, H, Bs = prob.host()
    assert rel_fro(W, g[f"W_it{it}"]) < 2e-4 and rel_fro(H, g[f"H_it{it}"]) < 2e-4
    for k in range(W.shape[1]):
        assert_same_top_ranking(W[:, k], g[f"W_it{it}"][:, k], what=("W column", k))
    prob.solver.scale()
    W, H, Bs = prob.host()
    assert rel_fro(W, g["W_scaled"]) < 2e-4 and rel_fro(H, g["H_scaled"]) < 2e-4
    # the public query on a model holding the GPU result, as fit() leaves it (main.py:143, 246-273)
    kw = dict(CASE_KW[name])
    model = ALPINE(device="cuda:0", **kw)
    blocks = [int(b) for b in g["blocks"]]
    cuts = np.cumsum([0] + blocks)
    keys = [f"cov{i}" for i in range(n_cov)]
    labels = {k: pd.Series([np.nan if na else str(v) for v, na in zip(g[f"labels{i}"], g[f"labels{i}_isna"])], dtype=object)
              for i, k in enumerate(keys)}
    model.covariate_keys = keys
    model.feature_names = [f"gene{j}" for j in range(W.shape[0])]
    model.fe = FeatureEncoders(keys)
    Y = model.fe.fit_transform(pd.DataFrame(labels))
    model.matrices = {"X": None, "Ys": [np.ascontiguousarray(y.T) for y in Y],
                      "Ws": [W[:, cuts[b]:cuts[b + 1]] for b in range(len(blocks))],
                      "Hs": [H[cuts[b]:cuts[b + 1]] for b in range(len(blocks))], "Bs": Bs}
    scores = model.get_covariate_gene_scores()
    for i, key in enumerate(keys):
        got, ref = scores[key].to_numpy(), g[f"gene_scores{i}"]
        assert got.shape == ref.shape
        assert rel_fro(got, ref) < 2e-4
        for c in range(ref.shape[1]):
            assert_same_top_ranking(got[:, c], ref[:, c], what=(key, c))
