"""Helpers for the `-m gpu` parity tests: golden fixture -> device-resident Solver (through the C ABI)."""
from __future__ import annotations

import numpy as np
import torch

from alpine_b200 import _native
from tests.helpers import CASE_KW


def to_dev_padded(a: np.ndarray, device) -> torch.Tensor:
    t = _native.padded_rows(a.shape[0], a.shape[1], device)
    t.copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)))
    return t


def csr_to_dev(X_cells_by_genes, device):
    """(indptr int64, indices int32, values fp32) device tensors of the CSR form (over cells) of a dense or scipy
    matrix."""
    import scipy.sparse as sp

    m = sp.csr_matrix(X_cells_by_genes, dtype=np.float32)
    m.sum_duplicates()
    return (torch.from_numpy(m.indptr.astype(np.int64)).to(device), torch.from_numpy(m.indices.astype(np.int32)).to(device),
            torch.from_numpy(m.data.astype(np.float32)).to(device))


class DeviceProblem:
    """X (cells x genes), Ys (c_i x n), W0 (G x K), H0 (K x n), Bs0 on the GPU, bound to a native Solver."""

    def __init__(self, X_cells_by_genes, Ys, W0, H0, Bs0, blocks, kw, device="cuda:0", sparse=False,
                 exchange_buffer=False):
        self.device = torch.device(device)
        n, G = X_cells_by_genes.shape
        self.X = None if sparse else to_dev_padded(X_cells_by_genes, self.device)
        self.Ys = [torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).to(self.device) for y in Ys]
        self.W = torch.from_numpy(np.ascontiguousarray(W0, dtype=np.float32)).to(self.device)
        self.H = to_dev_padded(H0, self.device)
        self.Bs = [torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32)).to(self.device) for b in Bs0]
        self.solver = _native.Solver(self.device, G, n, blocks, [y.shape[0] for y in Ys],
                                     kw.get("loss_type", "kl-divergence"))
        if sparse:
            self.solver.bind_csr(*csr_to_dev(X_cells_by_genes, self.device))
        else:
            self.solver.bind_dense(self.X)
        self.solver.bind_labels(self.Ys)
        self.solver.bind_factors(self.W, self.H, self.Bs)
        self.solver.set_hparams(kw.get("lam", []), kw.get("alpha_W", 0.0), kw.get("l1_ratio_W", 0.0),
                                kw.get("orth_W", 0.0), kw.get("eps", 1e-6))
        # exchange_buffer=True binds a caller-owned reduce buffer as the multi-GPU engine does: the numerator of the
        # W update then goes through it (alpine_mu_partials reduces the partial sums into it) instead of being summed
        # from the contraction's slots inside the update kernel; the block Gauss-Seidel sweep always uses it
        self.exchange_buffer = exchange_buffer
        if exchange_buffer:
            self.solver.reduce_buffer()

    def run(self, n_iter, on_iter=None, use_als=False):
        s = self.solver
        s.fit_begin(n_iter)
        for it in range(n_iter):
            s.mu_partials()
            if use_als:  # block Gauss-Seidel sweep (main.py:523-588)
                for b in range(s.n_blocks):
                    s.als_block(b)
                s.als_finish(it)
            else:
                s.mu_apply(it)
            if on_iter is not None:
                on_iter(it + 1)
        return s.losses(n_iter)

    def host(self):
        self.solver.sync_w()  # the updates keep W^T; the bound row-major W is refreshed on demand
        return self.W.cpu().numpy(), self.H.cpu().numpy(), [b.cpu().numpy() for b in self.Bs]


def problem_from_golden(name, g, device="cuda:0", sparse=False, exchange_buffer=False) -> DeviceProblem:
    n_cov = int(g["n_cov"])
    Ys = [np.ascontiguousarray(g[f"Y{i}_cells_by_cat"].T) for i in range(n_cov)]
    kw = dict(CASE_KW[name])
    kw.pop("use_als", None)
    return DeviceProblem(g["X_cells_by_genes"], Ys, g["W0"], g["H0"], [g[f"B0_{i}"] for i in range(n_cov)],
                         [int(b) for b in g["blocks"]], kw, device, sparse=sparse, exchange_buffer=exchange_buffer)
