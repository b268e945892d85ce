"""alpine_batch_gather / alpine_batch_scatter (the advanced-indexing gathers of main.py:593-595 and the scatter of
main.py:662) against torch indexing: bit-exact, including zero padding, repeated cells and a bad cell number."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from alpine_b200 import _native

pytestmark = pytest.mark.gpu


def _batch_context(dev, G, size, K, c_cov, dense=True):
    Xb = _native.padded_rows(size, G, dev) if dense else None
    Hb = _native.padded_rows(K, size, dev)
    Yb = [torch.empty((c, size), dtype=torch.float32, device=dev) for c in c_cov]
    W = torch.rand((G, K), device=dev)
    Bs = [torch.rand((c, 2), device=dev) for c in c_cov]
    s = _native.Solver(dev, G, size, [2] * len(c_cov) + [K - 2 * len(c_cov)], c_cov, "kl-divergence")
    if dense:
        s.bind_dense(Xb)
    s.bind_labels(Yb)
    s.bind_factors(W, Hb, Bs)
    s.set_hparams([1.0] * len(c_cov), 0.0, 0.0, 0.0, 1e-6)
    return s, Xb, Hb, Yb


@pytest.mark.parametrize("G,n_all,size,cnt", [(37, 500, 256, 200), (1030, 3000, 512, 512), (64, 90, 90, 1)])
def test_gather_and_scatter_equal_torch_indexing(G, n_all, size, cnt):
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(G + cnt)
    K, c_cov = 9, [3, 4]
    s, Xb, Hb, Yb = _batch_context(dev, G, size, K, c_cov)
    try:
        X_all = _native.padded_rows(n_all, G, dev)
        X_all.copy_(torch.rand((n_all, G), device=dev, generator=gen))
        H_all = _native.padded_rows(K, n_all, dev)
        H_all.copy_(torch.rand((K, n_all), device=dev, generator=gen))
        Ys_all = [torch.rand((c, n_all), device=dev, generator=gen) for c in c_cov]
        idx = torch.randint(0, n_all, (cnt,), device=dev, generator=gen)  # with repetition, like the weighted sampler
        for t in (Xb, Hb, *Yb):
            t.fill_(7.0)  # stale content of the previous batch
        s.batch_gather(X_all, H_all, Ys_all, idx)
        torch.cuda.synchronize(dev)
        assert torch.equal(Xb[:cnt], X_all[idx]) and not Xb[cnt:].any()
        assert torch.equal(Hb[:, :cnt], H_all[:, idx]) and not Hb[:, cnt:].any()
        for yb, y in zip(Yb, Ys_all):
            assert torch.equal(yb[:, :cnt], y[:, idx]) and not yb[:, cnt:].any()

        # the scatter: unique cells, so that torch's result is defined too
        uniq = torch.unique(idx)
        Hb[:, :len(uniq)] = torch.rand((K, len(uniq)), device=dev, generator=gen)
        want = H_all.clone()
        want[:, uniq] = Hb[:, :len(uniq)]
        s.batch_scatter(H_all, uniq)
        torch.cuda.synchronize(dev)
        assert torch.equal(H_all, want)
    finally:
        s.close()


def test_gather_reports_a_cell_number_outside_the_data():
    dev = torch.device("cuda:0")
    s, Xb, Hb, Yb = _batch_context(dev, 40, 256, 6, [3])
    try:
        X_all = _native.padded_rows(100, 40, dev).fill_(1.0)
        H_all = _native.padded_rows(6, 100, dev).fill_(1.0)
        Ys_all = [torch.ones((3, 100), device=dev)]
        idx = torch.tensor([5, 100, 7], device=dev)
        s.batch_gather(X_all, H_all, Ys_all, idx)
        s.batch_begin()
        with pytest.raises(_native.AlpineNativeError, match="idx\\[1\\] = 100"):
            s.losses(0)
        assert not Xb[1].any() and Xb[0].any() and Xb[2].any()  # the bad cell became padding, nothing was read outside
    finally:
        s.close()


def test_gather_for_a_csr_context_takes_factors_and_labels_only():
    dev = torch.device("cuda:0")
    from tests.gpu_utils import csr_to_dev

    s, _, Hb, Yb = _batch_context(dev, 50, 256, 6, [3], dense=False)
    try:
        rng = np.random.default_rng(0)
        s.bind_csr(*csr_to_dev(rng.random((256, 50)) * (rng.random((256, 50)) < 0.1), dev))  # the batch's own rows
        H_all = _native.padded_rows(6, 400, dev)
        H_all.copy_(torch.rand((6, 400), device=dev))
        Ys_all = [torch.rand((3, 400), device=dev)]
        idx = torch.arange(399, 199, -1, device=dev)
        s.batch_gather(None, H_all, Ys_all, idx)
        torch.cuda.synchronize(dev)
        assert torch.equal(Hb[:, :200], H_all[:, idx]) and not Hb[:, 200:].any()
        assert torch.equal(Yb[0][:, :200], Ys_all[0][:, idx])
    finally:
        s.close()
