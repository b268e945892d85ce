"""Torch restatement of the reference's simultaneous-update step, for the timing legs of bench.py and, on
``device="cuda"``, as the same-box fp32 trajectory the large-shape GPU parity tests compare against.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/alpine_oracle.py for the rules).  The reference runs its loop
with torch operators on ``device="cpu"`` (MKL GEMMs, torch's threading); the NumPy oracle computes the same numbers
but through NumPy's BLAS, which is about half as fast on these shapes.  A CPU baseline should not be slower than
the thing it stands for, so ``bench.py``'s ``cpu_baseline`` and ``--impl reference`` legs time THIS restatement:
the same operators, operator precedence, temporaries and batch gather as ``alpine/main.py:589-663`` (step) and
``main.py:726-753`` (loss), on torch CPU tensors with all host threads.  Pinned by tests/test_oracle_golden.py
against the same reference-generated fixtures as the NumPy oracle.  Every tensor it creates lives on the device of
its inputs, so the same code is the reference's ``device="cuda"`` arithmetic (cuBLAS fp32 GEMMs; callers switch
``allow_tf32`` off) when handed CUDA tensors: tests/test_gpu_bigshape.py and bench.py's ``gpu_torch_baseline``.
"""
from __future__ import annotations

from typing import List

import torch

from .alpine_oracle import KL, HyperParams


def block_slices(blocks: List[int]) -> List[slice]:
    out, s = [], 0
    for k in blocks:
        out.append(slice(s, s + k))
        s += k
    return out


@torch.no_grad()
def mu_step(X: torch.Tensor, Ys: List[torch.Tensor], W: torch.Tensor, H: torch.Tensor, Bs: List[torch.Tensor],
            blocks: List[int], hp: HyperParams, perm: torch.Tensor = None) -> None:
    """In place on W, H, Bs.  ``perm`` is the epoch's index vector (full batch: a permutation of all cells, which
    the reference draws and gathers with every iteration, main.py:502-521)."""
    sls = block_slices(blocks)
    n_cov = len(hp.n_covariate_components)
    eps = hp.eps
    if perm is None:  # natural order without the gather copies (the permutation only reorders fp32 sums)
        X_b, Ys_b, H_b = X, list(Ys), H
    else:
        X_b = X[:, perm]  # main.py:520
        Ys_b = [Y[:, perm] for Y in Ys]  # main.py:521
        H_b = H[:, perm]  # main.py:593-594
    K = W.shape[1]
    # === W === (main.py:592-612)
    numerator = (2 * X_b) @ H_b.T  # main.py:596
    orth = hp.orth_W * (torch.ones((K, K), dtype=W.dtype, device=W.device)
                        - torch.eye(K, dtype=W.dtype, device=W.device))  # main.py:474-484
    denominator = ((2 * W) @ H_b) @ H_b.T + ((1 - hp.l1_ratio_W) * hp.alpha_W) * W + W @ orth  # main.py:599-601
    denominator += hp.l1_ratio_W * hp.alpha_W * torch.ones_like(denominator)  # main.py:603
    denominator = torch.clamp(denominator, min=eps)  # main.py:604
    W *= numerator / denominator  # main.py:605
    # === B === (main.py:615-628), old H, old B
    for i in range(n_cov):
        Yb, Hb, B = Ys_b[i], H_b[sls[i]], Bs[i]
        if hp.loss_type == KL:
            num = (hp.lam[i] * (Yb / torch.clamp(B @ Hb, min=eps))) @ Hb.T
            den = (hp.lam[i] * torch.ones_like(Yb)) @ Hb.T
        else:
            num = (2 * Yb) @ Hb.T
            den = ((2 * B) @ Hb) @ Hb.T
        B *= num / torch.clamp(den, min=eps)
    # === H === (main.py:631-663), new W, new B, old H
    numerator = torch.zeros_like(H_b)
    denominator = torch.zeros_like(H_b)
    for i in range(n_cov):
        B = Bs[i]
        if hp.loss_type == KL:
            numerator[sls[i]] = (hp.lam[i] * B.T) @ (Ys_b[i] / torch.clamp(B @ H_b[sls[i]], min=eps))
            denominator[sls[i]] = (hp.lam[i] * B.T) @ torch.ones_like(Ys_b[i])
        else:
            numerator[sls[i]] = (2 * hp.lam[i] * B.T) @ Ys_b[i]
            denominator[sls[i]] = (2 * hp.lam[i] * B.T) @ (B @ H_b[sls[i]])
    numerator += (2 * W.T) @ X_b  # main.py:653
    denominator += (2 * W.T) @ (W @ H_b)  # main.py:654
    denominator = torch.clamp(denominator, min=eps)  # main.py:655
    if perm is None:
        H.copy_(H_b * (numerator / denominator))
    else:
        H[:, perm] = H_b * (numerator / denominator)  # main.py:656-663


@torch.no_grad()
def compute_loss(X, Ys, W, H, Bs, blocks, hp: HyperParams) -> List[float]:
    """[total, reconstruction, prediction...] as main.py:726-753 (fp32, torch.norm)."""
    sls = block_slices(blocks)
    recon = float(torch.norm(X - W @ H, p="fro") ** 2)
    pred = []
    for i, Y in enumerate(Ys):
        y_hat = Bs[i] @ H[sls[i]]
        if hp.loss_type == KL:
            y_hat = torch.clamp(y_hat, min=hp.eps)
            pred.append(float(torch.sum(Y * torch.log(torch.clamp(Y / y_hat, min=hp.eps)) - Y + y_hat)))
        else:
            pred.append(float(torch.norm(Y - y_hat, p="fro") ** 2))
    return [recon + sum(hp.lam[i] * p for i, p in enumerate(pred)), recon] + pred
