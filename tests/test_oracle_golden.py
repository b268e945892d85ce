"""Pin the NumPy oracle against trajectories produced by the unmodified reference.

The fixtures under tests/golden/ were written by oracle/gen_golden.py, which
runs the reference's own _initialize_matrices/_fit/_scale_matrices.  Tolerances:
the oracle and the reference (torch/MKL) differ only in fp32 summation order.
"""
import numpy as np
import pytest

from oracle import alpine_oracle as orc
from tests.helpers import (CASE_KW, assert_same_top_ranking, epoch_batches, golden_names, hp_of, inputs_of,
                           load_golden, rel_fro)

def full_batch_mu_names_for_torch():
    from tests.helpers import full_batch_mu_names

    return [n for n in full_batch_mu_names() if n not in ("kl_long200", "kl_scores2k")]  # no snapshots <= 10


TRAJ_TOL = 2e-6  # Frobenius-relative, per kept iteration (<= 10 iterations)
LONG_TOL = 5e-5  # 200 iterations of drift


@pytest.mark.parametrize("name", golden_names())
def test_step_trajectory_matches_reference(name):
    g = load_golden(name)
    hp = hp_of(name)
    X, Ys, st = inputs_of(g)
    use_als = CASE_KW[name].get("use_als", False)
    kept = set(int(i) for i in g["kept_iters"])
    n_iter = int(max(kept))
    tol = LONG_TOL if n_iter > 20 else TRAJ_TOL
    for it in range(1, n_iter + 1):
        for idx in epoch_batches(g, it):
            (orc.als_step if use_als else orc.mu_step)(X, Ys, st, hp, idx=idx)
        if it in kept:
            assert rel_fro(st.W, g[f"W_it{it}"]) < tol, (name, it, "W")
            assert rel_fro(st.H, g[f"H_it{it}"]) < tol, (name, it, "H")
            for i in range(len(Ys)):
                assert rel_fro(st.Bs[i], g[f"B{i}_it{it}"]) < tol, (name, it, f"B{i}")


@pytest.mark.parametrize("name", [n for n in golden_names() if n != "kl_long200"])
def test_loss_matches_reference(name):
    g = load_golden(name)
    hp = hp_of(name)
    X, Ys, st = inputs_of(g)
    it = int(g["kept_iters"][-1])
    st.W, st.H = g[f"W_it{it}"].copy(), g[f"H_it{it}"].copy()
    st.Bs = [g[f"B{i}_it{it}"].copy() for i in range(len(Ys))]
    ref = g["loss_history_ref_fp32"][it - 1]
    got32 = orc.compute_loss(X, Ys, st, hp)
    got64 = orc.compute_loss(X, Ys, st, hp, dtype=np.float64)
    # fp32 torch.norm vs NumPy summation order: the gap grows with the element count
    np.testing.assert_allclose(got32, ref, rtol=2e-5 if X.size < 200_000 else 1e-4)
    np.testing.assert_allclose(got64[1], float(g["final_recon_fp64"]), rtol=1e-10)
    np.testing.assert_allclose(got64[:2], ref[:2], rtol=1e-4)
    # the KL terms y*log(y/yhat) - y + yhat cancel; fp32 leaves an absolute error ~ eps32 * sum(y)
    np.testing.assert_allclose(got64[2:], ref[2:], rtol=1e-4, atol=1e-6 * X.shape[1])


@pytest.mark.parametrize("name", golden_names())
def test_scale_transform_scores_match_reference(name):
    g = load_golden(name)
    hp = hp_of(name)
    X, Ys, st = inputs_of(g)
    it = int(g["kept_iters"][-1])
    st.W, st.H = g[f"W_it{it}"].copy(), g[f"H_it{it}"].copy()
    st.Bs = [g[f"B{i}_it{it}"].copy() for i in range(len(Ys))]
    orc.scale_matrices(st, hp)
    assert rel_fro(st.W, g["W_scaled"]) < 1e-6
    assert rel_fro(st.H, g["H_scaled"]) < 1e-6
    for i in range(len(Ys)):
        assert rel_fro(st.Bs[i], g[f"B{i}_scaled"]) < 1e-6
    scores = orc.covariate_gene_scores(st.Ws(), st.Hs(), Ys)
    for i, s in enumerate(scores):
        assert rel_fro(s, g[f"gene_scores{i}"]) < 1e-6
    Ht = orc.transform_loop(X, g["W_scaled"], g["Ht0"], 5, hp.eps)
    assert rel_fro(Ht, g["Ht_5"]) < 2e-6


@pytest.mark.parametrize("name", [n for n in golden_names()])
def test_one_hot_matches_reference_encoder(name):
    g = load_golden(name)
    for i in range(int(g["n_cov"])):
        labels = [np.nan if na else str(v) for v, na in zip(g[f"labels{i}"], g[f"labels{i}_isna"])]
        Y, cats = orc.one_hot(labels)
        np.testing.assert_array_equal(Y, g[f"Y{i}_cells_by_cat"])
        # sklearn names the columns "<key>_<category>"
        assert [c.split("_", 1)[1] for c in g[f"cats{i}"]] == cats


def test_long_run_top100_rankings_match_reference():
    """Gene rankings per component after 200 iterations (north_star parity criterion)."""
    g = load_golden("kl_long200")
    hp = hp_of("kl_long200")
    X, Ys, st = inputs_of(g)
    for _ in range(200):
        orc.mu_step(X, Ys, st, hp)
    Wref = g["W_it200"]
    for k in range(st.W.shape[1]):
        assert_same_top_ranking(st.W[:, k], Wref[:, k], what=("W column", k))


@pytest.mark.parametrize("name", full_batch_mu_names_for_torch())
def test_torch_port_matches_reference(name):
    """The torch-CPU restatement that bench.py times as the CPU baseline, on the reference's own trajectories."""
    import torch

    from oracle import torch_port as tp

    torch.set_num_threads(1)
    torch.manual_seed(0)  # the permutations below only reorder fp32 sums; seeded so that the test is deterministic
    g = load_golden(name)
    hp = hp_of(name)
    X, Ys, st = inputs_of(g)
    Xt = torch.from_numpy(np.ascontiguousarray(X))
    Yt = [torch.from_numpy(y.copy()) for y in Ys]
    W, H = torch.from_numpy(st.W.copy()), torch.from_numpy(st.H.copy())
    Bs = [torch.from_numpy(b.copy()) for b in st.Bs]
    kept = [int(i) for i in g["kept_iters"] if int(i) <= 10]
    for it in range(1, max(kept) + 1):
        tp.mu_step(Xt, Yt, W, H, Bs, st.blocks, hp, perm=torch.randperm(X.shape[1]))
        if it in kept:
            assert rel_fro(W.numpy(), g[f"W_it{it}"]) < TRAJ_TOL and rel_fro(H.numpy(), g[f"H_it{it}"]) < TRAJ_TOL
            for i in range(len(Ys)):
                assert rel_fro(Bs[i].numpy(), g[f"B{i}_it{it}"]) < 3 * TRAJ_TOL  # (a dozen entries: less averaging)
    if max(kept) == int(g["kept_iters"][-1]):
        loss = tp.compute_loss(Xt, Yt, W, H, Bs, st.blocks, hp)
        np.testing.assert_allclose(loss[:2], g["loss_history_ref_fp32"][max(kept) - 1][:2], rtol=5e-5)
