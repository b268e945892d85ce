"""GPU checks at BASELINE.json's full sizes through size-independent properties (the oracle would take minutes
there): linearity of the contractions in the cells, cell-block additivity (what the multi-GPU sharding relies on),
CSR == dense, monotone objective, non-negativity, and scaling invariance."""
import numpy as np
import pytest
import torch

from tests.helpers import rel_fro

pytestmark = pytest.mark.gpu


def _device_problem(n, G, blocks, cats, seed=0, density=None):
    from alpine_b200 import _native

    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    K = sum(blocks)
    X = _native.padded_rows(n, G, dev)
    step = max(1, (1 << 27) // G)
    for r0 in range(0, n, step):
        r1 = min(n, r0 + step)
        blk = torch.rand((r1 - r0, G), device=dev, generator=g).pow_(3.0)
        if density is not None:
            blk = torch.where(torch.rand((r1 - r0, G), device=dev, generator=g) < density,
                              torch.floor(blk * 8.0) + 1.0, torch.zeros((), device=dev))
        X[r0:r1] = blk
    W = torch.rand((G, K), device=dev, generator=g).clamp_(min=1e-6)
    H = _native.padded_rows(K, n, dev)
    H.copy_(torch.rand((K, n), device=dev, generator=g).clamp_(min=1e-6))
    Ys = [torch.nn.functional.one_hot(torch.randint(0, c, (n,), device=dev, generator=g), c).T.contiguous().float()
          for c in cats]
    Bs = [torch.rand((c, k), device=dev, generator=g).clamp_(min=1e-6).contiguous() for c, k in zip(cats, blocks)]
    return X, Ys, W, H, Bs


def _solver(X, Ys, W, H, Bs, blocks, cats, lam, sparse=False):
    from alpine_b200 import _native

    n, G = X.shape
    s = _native.Solver(X.device, G, n, blocks, cats)
    if sparse:
        csr = X.contiguous().to_sparse_csr()
        s.bind_csr(csr.crow_indices().to(torch.int64), csr.col_indices().to(torch.int32), csr.values().contiguous())
    else:
        s.bind_dense(X)
    s.bind_labels(Ys)
    s.bind_factors(W, H, Bs)
    s.set_hparams(lam, 0.5, 0.5, 0.2, 1e-6)
    return s


def test_cfg3_contractions_are_additive_over_cell_blocks():
    """20,000 genes x 100,000 cells, K = 100: X H^T over all cells == sum over two cell blocks (the all-reduce the
    sharded engine performs), and W^T X of a block == the block of W^T X."""
    n, G, blocks = 100000, 20000, [5, 5, 90]
    X, Ys, W, H, Bs = _device_problem(n, G, blocks, [3, 4])
    full = _solver(X, Ys, W, H, Bs, blocks, [3, 4], [1e3, 1e3])
    xh, wx = full.xh_product().double(), full.wx_product().clone()
    full.close()
    half = n // 2 + 37  # ragged split, not a multiple of the 256-row tile
    acc = torch.zeros_like(xh)
    for lo, hi in ((0, half), (half, n)):
        from alpine_b200 import _native

        Hb = _native.padded_rows(sum(blocks), hi - lo, X.device)
        Hb.copy_(H[:, lo:hi])
        part = _solver(X[lo:hi], [y[:, lo:hi].contiguous() for y in Ys], W, Hb, Bs, blocks, [3, 4], [1e3, 1e3])
        acc += part.xh_product().double()
        assert rel_fro(part.wx_product().cpu().numpy(), wx[:, lo:hi].cpu().numpy()) < 1e-6
        part.close()
    assert rel_fro(acc.cpu().numpy(), xh.cpu().numpy()) < 1e-6
    # spot check of 64 genes against fp64 on the full reduction length
    ref = H.double() @ X[:, :64].double()
    assert rel_fro(xh[:, :64].cpu().numpy(), ref.cpu().numpy()) < 3e-6


def test_cfg2_fit_properties_dense_and_csr():
    """5,000 HVG x 50,000 cells, 30 + [5, 5] components, lam = [1e3, 1e3] (BASELINE configs[1]): monotone objective,
    non-negative factors, CSR == dense, and W H invariant under the post-fit scaling."""
    n, G, blocks, cats, lam = 50000, 5000, [5, 5, 30], [3, 4], [1e3, 1e3]
    X, Ys, W0, H0, B0 = _device_problem(n, G, blocks, cats, seed=2, density=0.15)
    outs = []
    for sparse in (False, True):
        W, H, Bs = W0.clone(), H0.clone(), [b.clone() for b in B0]
        from alpine_b200 import _native

        Hp = _native.padded_rows(sum(blocks), n, X.device)
        Hp.copy_(H)
        s = _solver(X, Ys, W, Hp, Bs, blocks, cats, lam, sparse=sparse)
        n_iter = 30
        s.fit_begin(n_iter)
        for it in range(n_iter):
            s.mu_partials()
            s.mu_apply(it)
        xn, rows = s.losses(n_iter)
        recon = xn - 2.0 * rows[:, 0] + rows[:, 1]
        total = recon + sum(l * rows[:, 2 + i] for i, l in enumerate(lam))
        assert np.all(np.isfinite(total)) and np.all(recon > 0)
        assert np.all(np.diff(total) <= 1e-6 * total[:-1]), "the MU objective must not increase"
        assert float(W.min()) >= 0 and float(Hp.min()) >= 0 and all(float(b.min()) >= 0 for b in Bs)
        # scaling (main.py:772-781): column sums of W become 1, W H is unchanged
        probe = (W[:256] @ Hp[:, :512]).double()
        s.scale()
        torch.cuda.synchronize()
        np.testing.assert_allclose(W.sum(dim=0).cpu().numpy(), 1.0, rtol=1e-5)
        assert rel_fro((W[:256] @ Hp[:, :512]).double().cpu().numpy(), probe.cpu().numpy()) < 1e-5
        outs.append((W.cpu().numpy(), Hp.cpu().numpy(), total))
        s.close()
    assert rel_fro(outs[1][0], outs[0][0]) < 1e-6 and rel_fro(outs[1][1], outs[0][1]) < 1e-6
    np.testing.assert_allclose(outs[1][2], outs[0][2], rtol=1e-9)


def test_initial_factors_equal_default_generator_stream():
    """_initialize_matrices draws from an explicit per-device generator (thread safety of the fold scheduler); the
    values must equal what the reference's global seeding + default generator give (main.py:440-470)."""
    from alpine_b200 import ALPINE

    model = ALPINE(n_components=7, n_covariate_components=[3], lam=[1.0], device="cuda:0", random_state=11)
    n, G = 300, 120
    X = np.random.default_rng(0).random((n, G), dtype=np.float32).T
    Y = [np.eye(3, dtype=np.float32)[np.random.default_rng(1).integers(0, 3, n)]]
    m = model._initialize_matrices(X, Y)
    torch.manual_seed(11)
    torch.cuda.manual_seed(11)
    dev = torch.device("cuda:0")
    Ws = [torch.rand((G, k), dtype=torch.float32, device=dev).clamp(min=1e-6) for k in (3, 7)]
    Hs = [torch.rand((k, n), dtype=torch.float32, device=dev).clamp(min=1e-6) for k in (3, 7)]
    B = torch.rand((3, 3), dtype=torch.float32, device=dev).clamp(min=1e-6)
    assert torch.equal(m.W, torch.cat(Ws, 1)) and torch.equal(m.H, torch.cat(Hs, 0)) and torch.equal(m.Bs[0], B)
