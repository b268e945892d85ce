"""CPU oracle for ALPINE's covariate-guided multiplicative-update NMF loop.

TEST INFRASTRUCTURE ONLY.  This module is a NumPy restatement of the reference
algorithm (ylaboratory/ALPINE, ``alpine/main.py``) and exists so that the CUDA
path can be checked against it.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it;
the product package ``alpine_b200`` never does.

Parity status: the reference ships no tests or golden vectors
(``tests/__init__.py`` is empty), so this restatement is pinned against
outputs of the reference itself: ``oracle/gen_golden.py`` imports the
unmodified reference from ``/root/reference`` (through the import shim in
``oracle/ref_shim.py``), runs ``ALPINE._initialize_matrices`` / ``_fit`` /
``_scale_matrices`` / the ``_transform`` loop on seeded inputs and commits the
trajectories under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks
every function here against those fixtures.

Every function cites the reference lines it follows.  The arithmetic is kept
*literal* (same operator precedence, same materialised temporaries) so that
fp32 round-off is of the same kind as the reference's; the Gram/trace
reformulation used by the CUDA kernels is deliberately NOT used here.

Conventions (reference ``main.py:28-43``): X is genes x cells (G, n); Ys[i] is
c_i x n one-hot (all-zero column for a NaN label); W = cat(Ws, 1) is G x K;
H = cat(Hs, 0) is K x n; Bs[i] is c_i x k_i.  Blocks are ordered covariates
first, unguided block last (``main.py:79``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

KL = "kl-divergence"
FROB = "frobenius"


@dataclass
class HyperParams:
    """Constructor arguments of the reference that the loop reads (main.py:47-73)."""

    n_components: int
    n_covariate_components: List[int]
    lam: List[float]
    orth_W: float = 0.0
    alpha_W: float = 0.0
    l1_ratio_W: float = 0.0
    loss_type: str = KL
    eps: float = 1e-6

    @property
    def n_all_components(self) -> List[int]:  # main.py:79
        return list(self.n_covariate_components) + [self.n_components]

    @property
    def total_components(self) -> int:  # main.py:80
        return int(sum(self.n_all_components))


@dataclass
class State:
    """Concatenated factor matrices plus block bookkeeping."""

    W: np.ndarray  # G x K
    H: np.ndarray  # K x n
    Bs: List[np.ndarray]  # c_i x k_i
    blocks: List[int] = field(default_factory=list)  # k_i per block, unguided last

    def copy(self) -> "State":
        return State(self.W.copy(), self.H.copy(), [b.copy() for b in self.Bs], list(self.blocks))

    def block_slices(self) -> List[slice]:
        out, s = [], 0
        for k in self.blocks:
            out.append(slice(s, s + k))
            s += k
        return out

    def Ws(self) -> List[np.ndarray]:
        return [self.W[:, sl] for sl in self.block_slices()]

    def Hs(self) -> List[np.ndarray]:
        return [self.H[sl, :] for sl in self.block_slices()]


# ----------------------------------------------------------------------------
# encoder.py:17-38  one-hot encoding with NaN -> all-zero row
# ----------------------------------------------------------------------------
def one_hot(labels: Sequence, dtype=np.float32) -> Tuple[np.ndarray, List[str]]:
    """cells x categories dummy matrix (encoder.py:27-37).

    Categories are the sorted unique non-null labels (what sklearn's
    OneHotEncoder does); null labels give an all-zero row (encoder.py:32-35).
    """
    lab = np.asarray(labels, dtype=object)
    isna = np.array([(v is None) or (isinstance(v, float) and np.isnan(v)) for v in lab])
    cats = sorted(set(lab[~isna].tolist()))
    out = np.zeros((len(lab), len(cats)), dtype=dtype)
    index = {c: i for i, c in enumerate(cats)}
    for r, v in enumerate(lab):
        if not isna[r]:
            out[r, index[v]] = 1
    return out, [str(c) for c in cats]


# ----------------------------------------------------------------------------
# main.py:474-484  orthogonality matrix
# ----------------------------------------------------------------------------
def orth_matrix(size: int, orth_W: float, dtype) -> np.ndarray:
    m = np.ones((size, size), dtype=dtype)
    m -= np.eye(size, dtype=dtype)
    return (dtype(orth_W) * m).astype(dtype)


# ----------------------------------------------------------------------------
# main.py:589-663  one non-ALS batch step (full batch unless idx is given)
# ----------------------------------------------------------------------------
def mu_step(
    X: np.ndarray,
    Ys: List[np.ndarray],
    st: State,
    hp: HyperParams,
    idx: Optional[np.ndarray] = None,
    literal_cost: bool = False,
) -> None:
    """In-place W, B, H multiplicative update.

    ``idx`` selects the batch columns (main.py:520-521, 593); ``None`` means the
    full batch in natural order, which differs from the reference's randperm
    only in floating-point summation order.  ``literal_cost=True`` additionally
    performs the reference's gather copy of X (main.py:520) even for the
    identity permutation so that CPU timings include it.
    """
    dt = st.W.dtype.type
    eps = dt(hp.eps)
    sls = st.block_slices()
    n_cov = len(hp.n_covariate_components)

    if idx is None:
        if literal_cost:
            idx_full = np.arange(X.shape[1])
            X_b = X[:, idx_full]
        else:
            X_b = X
        Ys_b = Ys
        H_b = st.H.copy()  # torch.cat copies (main.py:594)
    else:
        X_b = X[:, idx]
        Ys_b = [Y[:, idx] for Y in Ys]
        H_b = st.H[:, idx]  # fancy indexing copies
    Hs_b = [H_b[sl, :] for sl in sls]

    # === Update W === (main.py:592-612)
    W = st.W
    numerator = (dt(2) * X_b) @ H_b.T  # main.py:596  (2*X)@H^T
    om = orth_matrix(W.shape[1], hp.orth_W, dt)  # main.py:597
    denominator = (
        ((dt(2) * W) @ H_b) @ H_b.T  # main.py:599
        + dt((1 - hp.l1_ratio_W) * hp.alpha_W) * W  # main.py:600
        + W @ om  # main.py:601
    )
    denominator += dt(hp.l1_ratio_W * hp.alpha_W) * np.ones_like(denominator)  # main.py:603
    denominator = np.maximum(denominator, eps)  # main.py:604
    W *= numerator / denominator  # main.py:605

    # === Update Bs === (main.py:615-628), old H, old B
    for i in range(n_cov):
        Yb, Hb, B = Ys_b[i], Hs_b[i], st.Bs[i]
        if hp.loss_type == KL:
            num = (dt(hp.lam[i]) * (Yb / np.maximum(B @ Hb, eps))) @ Hb.T  # main.py:618-622
            den = (dt(hp.lam[i]) * np.ones_like(Yb)) @ Hb.T  # main.py:623
        else:
            num = (dt(2) * Yb) @ Hb.T  # main.py:625
            den = ((dt(2) * B) @ Hb) @ Hb.T  # main.py:626
        den = np.maximum(den, eps)  # main.py:627
        B *= num / den  # main.py:628

    # === Update H === (main.py:631-663), new W, new B, old H
    numerator = np.zeros_like(H_b)
    denominator = np.zeros_like(H_b)
    for i in range(n_cov):
        sl, B = sls[i], st.Bs[i]
        if hp.loss_type == KL:
            g_num = (dt(hp.lam[i]) * B.T) @ (Ys_b[i] / np.maximum(B @ Hs_b[i], eps))  # main.py:640-643
            g_den = (dt(hp.lam[i]) * B.T) @ np.ones_like(Ys_b[i])  # main.py:644
        else:
            g_num = (dt(2 * hp.lam[i]) * B.T) @ Ys_b[i]  # main.py:646
            g_den = (dt(2 * hp.lam[i]) * B.T) @ (B @ Hs_b[i])  # main.py:647
        numerator[sl] = g_num
        denominator[sl] = g_den
    numerator += (dt(2) * W.T) @ X_b  # main.py:653
    denominator += (dt(2) * W.T) @ (W @ H_b)  # main.py:654
    denominator = np.maximum(denominator, eps)  # main.py:655
    H_b = H_b * (numerator / denominator)  # main.py:656
    if idx is None:
        st.H[...] = H_b  # main.py:659-663
    else:
        st.H[:, idx] = H_b  # duplicate indices: last write wins, as torch index_put


# ----------------------------------------------------------------------------
# main.py:523-588  ALS / block Gauss-Seidel step
# ----------------------------------------------------------------------------
def als_step(X, Ys, st: State, hp: HyperParams, idx: Optional[np.ndarray] = None) -> None:
    dt = st.W.dtype.type
    eps = dt(hp.eps)
    sls = st.block_slices()
    n_cov = len(hp.n_covariate_components)
    cols = slice(None) if idx is None else idx
    X_b = X if idx is None else X[:, idx]
    Ys_b = Ys if idx is None else [Y[:, idx] for Y in Ys]
    for b, sl in enumerate(sls):
        H_cat = st.H[:, cols].copy()
        H_blk = H_cat[sl, :]
        W_blk = st.W[:, sl]
        W_cat = st.W.copy()
        numerator = (dt(2) * X_b) @ H_blk.T  # main.py:533
        k = W_blk.shape[1]
        denominator = (
            ((dt(2) * W_cat) @ H_cat) @ H_blk.T  # main.py:539
            + (dt((1 - hp.l1_ratio_W) * hp.alpha_W) * W_blk) @ np.eye(k, dtype=dt)  # main.py:540
            + W_blk @ orth_matrix(k, hp.orth_W, dt)  # main.py:541
        )
        denominator += dt(hp.l1_ratio_W * hp.alpha_W) * np.ones_like(denominator)  # main.py:543
        denominator = np.maximum(denominator, eps)
        st.W[:, sl] = W_blk * (numerator / denominator)  # main.py:545
        if b < n_cov:  # main.py:548-562
            Yb, B = Ys_b[b], st.Bs[b]
            if hp.loss_type == KL:
                num = (dt(hp.lam[b]) * (Yb / np.maximum(B @ H_blk, eps))) @ H_blk.T
                den = (dt(hp.lam[b]) * np.ones_like(Yb)) @ H_blk.T
            else:
                num = (dt(2) * Yb) @ H_blk.T
                den = ((dt(2) * B) @ H_blk) @ H_blk.T
            den = np.maximum(den, eps)
            B *= num / den
        W_blk = st.W[:, sl]  # main.py:565
        W_cat = st.W
        u_num = (dt(2) * W_blk.T) @ X_b  # main.py:567
        u_den = (dt(2) * W_blk.T) @ (W_cat @ H_cat)  # main.py:568
        if b < n_cov:  # main.py:570-585
            Yb, B = Ys_b[b], st.Bs[b]
            if hp.loss_type == KL:
                g_num = (dt(hp.lam[b]) * B.T) @ (Yb / np.maximum(B @ H_blk, eps))
                g_den = (dt(hp.lam[b]) * B.T) @ np.ones_like(Yb)
            else:
                g_num = (dt(2 * hp.lam[b]) * B.T) @ Yb
                g_den = (dt(2 * hp.lam[b]) * B.T) @ (B @ H_blk)
            num = u_num + g_num
            den = np.maximum(u_den + g_den, eps)
        else:  # main.py:586-588
            num = u_num
            den = np.maximum(u_den, eps)
        new = H_blk * (num / den)
        if idx is None:
            st.H[sl, :] = new
        else:
            st.H[sl, idx] = new


# ----------------------------------------------------------------------------
# main.py:726-753  per-iteration loss
# ----------------------------------------------------------------------------
def kl_divergence(y, y_hat, eps):
    y_hat = np.maximum(y_hat, eps)  # main.py:728
    return float(np.sum(y * np.log(np.maximum(y / y_hat, eps)) - y + y_hat))  # main.py:729-731


def compute_loss(X, Ys, st: State, hp: HyperParams, dtype=None) -> List[float]:
    """[total, recon, pred_0, ...] (main.py:750-753).

    ``dtype=np.float64`` re-evaluates the same formula in double precision from
    the given (fp32) factors; this is the loss-parity target because the
    reference's own fp32 ``torch.norm`` on CPU is ~6e-4 off.
    """
    dt = (dtype or st.W.dtype.type)
    X_, W_, H_ = X.astype(dt, copy=False), st.W.astype(dt, copy=False), st.H.astype(dt, copy=False)
    eps = dt(hp.eps)
    R = X_ - W_ @ H_
    recon = float(np.sqrt(np.sum(R * R, dtype=dt)) ** 2)  # torch.norm(...)**2, main.py:736
    pred = []
    for i, sl in enumerate(st.block_slices()[: len(Ys)]):
        Y_, B_ = Ys[i].astype(dt, copy=False), st.Bs[i].astype(dt, copy=False)
        if hp.loss_type == KL:
            pred.append(kl_divergence(Y_, B_ @ H_[sl], eps))  # main.py:741-743
        else:
            D = Y_ - B_ @ H_[sl]
            pred.append(float(np.sqrt(np.sum(D * D, dtype=dt)) ** 2))  # main.py:745-748
    total = recon + sum(hp.lam[i] * p for i, p in enumerate(pred))
    return [total, recon] + pred


# ----------------------------------------------------------------------------
# main.py:772-781  post-fit scaling
# ----------------------------------------------------------------------------
def scale_matrices(st: State, hp: HyperParams) -> None:
    n_cov = len(hp.n_covariate_components)
    for i, sl in enumerate(st.block_slices()):
        s = st.W[:, sl].sum(axis=0, dtype=st.W.dtype)  # main.py:776
        st.W[:, sl] = st.W[:, sl] / s  # main.py:777
        st.H[sl, :] = st.H[sl, :] * s[:, None]  # main.py:778
        if i < n_cov:
            st.Bs[i] = st.Bs[i] / s  # main.py:780-781


# ----------------------------------------------------------------------------
# main.py:705-709  H-only transform loop
# ----------------------------------------------------------------------------
def transform_loop(X, W, H0, n_iter: int, eps: float) -> np.ndarray:
    dt = W.dtype.type
    H = H0.copy()
    for _ in range(n_iter):
        numerator = (dt(2) * W.T) @ X  # main.py:706
        denominator = (dt(2) * W.T) @ (W @ H)  # main.py:707
        denominator = np.maximum(denominator, dt(eps))  # main.py:708
        H *= numerator / denominator  # main.py:709
    return H


# ----------------------------------------------------------------------------
# main.py:246-273  covariate gene scores
# ----------------------------------------------------------------------------
def covariate_gene_scores(Ws, Hs, Ys) -> List[np.ndarray]:
    out = []
    for W, H, Y in zip(Ws, Hs, Ys):
        HY = H @ Y.T / Y.sum(axis=1)  # main.py:260
        out.append(W @ HY)  # main.py:261
    return out


# ----------------------------------------------------------------------------
# main.py:486-676  driver: epochs of full-batch (or minibatch) steps + loss
# ----------------------------------------------------------------------------
def fit_loop(
    X,
    Ys,
    st: State,
    hp: HyperParams,
    max_iter: int,
    use_als: bool = False,
    record: bool = False,
    batch_perms: Optional[List[np.ndarray]] = None,
    batch_size: Optional[int] = None,
    literal_cost: bool = False,
):
    """Run ``max_iter`` epochs; returns (loss_history, trajectory).

    ``batch_perms[it]`` is the epoch index vector (sampling.py:6-16); with
    ``None`` each epoch is one full batch in natural order.
    """
    hist, traj = [], []
    n = X.shape[1]
    for it in range(max_iter):
        if batch_perms is None:
            if use_als:
                als_step(X, Ys, st, hp)
            else:
                mu_step(X, Ys, st, hp, literal_cost=literal_cost)
        else:
            perm = batch_perms[it]
            bs = batch_size or n
            for b0 in range(0, len(perm), bs):  # sampling.py:58-71
                idx = perm[b0 : b0 + bs]
                (als_step if use_als else mu_step)(X, Ys, st, hp, idx=idx)
        hist.append(compute_loss(X, Ys, st, hp))  # main.py:666
        if record:
            traj.append(st.copy())
    return hist, traj


# ----------------------------------------------------------------------------
# main.py:755-770  max_iter auto-detection (Kneedle elbow of log10 recon loss)
#
# PARITY UNPINNED: ``kneed`` is a third-party dependency of the reference
# (pyproject.toml:14-25, unpinned version) that is absent from /root/reference
# and from this image, so this restatement of kneed 0.8.x ``KneeLocator(x, y,
# S=1.0, curve="convex", direction="decreasing", interp_method="polynomial",
# polynomial_degree=2, online=False)`` (Satopaa et al. 2011) cannot be checked
# against the library here.  It is anchored on the reference's call site only.
# ----------------------------------------------------------------------------
def kneedle_elbow(x, y, S: float = 1.0, degree: int = 2) -> Optional[float]:
    from scipy.signal import argrelextrema

    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    Ds_y = np.poly1d(np.polyfit(x, y, degree))(x)  # step 1: smooth
    xn = (x - x.min()) / (x.max() - x.min())  # step 2: unit square
    yn = (Ds_y - Ds_y.min()) / (Ds_y.max() - Ds_y.min())
    yn = yn.max() - yn  # step 3: decreasing+convex elbow -> knee
    y_diff = yn - xn
    max_idx = argrelextrema(y_diff, np.greater_equal)[0]  # step 4
    min_idx = argrelextrema(y_diff, np.less_equal)[0]
    if max_idx.size == 0:
        return None
    Tmx = y_diff[max_idx] - S * np.abs(np.diff(xn).mean())  # step 5
    threshold, threshold_index, max_i = 0.0, 0, 0
    for i in range(len(xn)):  # step 6
        if i < max_idx[0]:
            continue
        if xn[i] == 1.0:
            break
        if (max_idx == i).any():
            threshold = Tmx[max_i]
            threshold_index = i
            max_i += 1
        if (min_idx == i).any():
            threshold = 0.0
        if y_diff[i + 1] < threshold:
            return float(x[threshold_index])
    return None
