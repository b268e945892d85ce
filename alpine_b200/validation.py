"""Argument validation with the reference's observable behaviour (exception type and message).

Reference: ``ALPINE._validate_init_args`` (main.py:322-381), ``_validate_fit_args`` (main.py:383-434), the checks
at the head of ``transform`` (main.py:155-164) and the ComponentOptimizer validators (optimization.py:512-604).
Quirks that are kept on purpose: ``lam`` entries and ``alpha_W`` / ``orth_W`` / ``l1_ratio_W`` /
``eps`` must be ``float`` instances (an ``int`` is rejected); ``adata.X`` must be a dense ``np.ndarray`` (or, as an
extension, a ``scipy.sparse`` matrix, which selects the CSR path);
covariate columns must have dtype kind ``'O'``; the ``batch_size`` / ``max_iter`` checks can never fire because of
how the reference chains ``and`` (main.py:420-428) and are therefore omitted.
"""
from __future__ import annotations

from typing import Any, Callable, Sequence, Tuple

import numpy as np

from .utils.anndata_compat import AnnData, is_sparse

LOSS_TYPES = ["kl-divergence", "frobenius"]


def _require(ok: bool, exc: type, msg: str) -> None:
    if not ok:
        raise exc(msg)


def _nonneg_float(value: Any, msg: str) -> None:
    _require(isinstance(value, float) and value >= 0, ValueError, msg)


def _all_non_negative(X: np.ndarray) -> bool:
    """``np.all(X >= 0)`` (main.py:399) without the boolean temporary: a multi-threaded min-reduction; NaN fails
    both forms."""
    if X.size == 0:
        return True
    try:
        import torch

        if X.dtype in (np.float32, np.float64) and X.flags.c_contiguous:
            return bool(torch.from_numpy(X).min().item() >= 0)
    except Exception:
        pass
    return bool(np.all(X >= 0))


def check_model_args(m) -> None:
    """main.py:322-381, in the reference's order."""
    _require(m.n_components > 0, ValueError, "n_components must be greater than 0.")
    _require(isinstance(m.n_covariate_components, list), TypeError, "n_covariate_components must be a list.")
    for n in m.n_covariate_components:
        _require(isinstance(n, int) and n >= 0, ValueError,
                 "Each element in n_covariate_components must be a non-negative integer.")
    _require(isinstance(m.lam, list), TypeError, "lam must be in a list.")
    for v in m.lam:
        _nonneg_float(v, "Each element in lam must be a non-negative float.")
    _nonneg_float(m.alpha_W, "alpha_W must be a non-negative float.")
    _nonneg_float(m.orth_W, "orth_W must be a non-negative float.")
    _require(isinstance(m.l1_ratio_W, float) and 0 <= m.l1_ratio_W <= 1, ValueError,
             "l1_ratio_W must be a float between 0 and 1.")
    _require(isinstance(m.scale_needed, bool), TypeError, "scale_needed must be a boolean.")
    _require(isinstance(m.loss_type, str), TypeError, "loss_type must be a string.")
    _require(m.loss_type in LOSS_TYPES, ValueError, f"loss_type must be one of {LOSS_TYPES}.")
    _nonneg_float(m.eps, "eps must be a non-negative float.")
    _require(isinstance(m.random_state, int) and m.random_state >= 0, ValueError,
             "random_state must be a non-negative integer.")


NONNEG_MSG = "All elements in adata.X must be non-negative."


def check_fit_args(m, adata, covariate_keys, batch_size, max_iter, sampling_method, verbose,
                   defer_nonneg: bool = False) -> bool:
    """main.py:383-434.  Returns True when the (expensive) non-negativity scan of a dense ``adata.X`` was deferred:
    the caller then checks the minimum on the device after the upload it has to do anyway and raises the same
    ``ValueError``.  The order of errors stays the reference's: if a later check fails, the scan is run on the host
    first, because the reference would have reported a negative X before that later error."""
    deferred = False
    try:
        deferred = _check_fit_args(m, adata, covariate_keys, batch_size, max_iter, sampling_method, verbose, defer_nonneg)
    except (TypeError, ValueError):
        if _check_fit_args.last_deferred:
            _require(_all_non_negative(adata.X), ValueError, NONNEG_MSG)
        raise
    return deferred


def _check_fit_args(m, adata, covariate_keys, batch_size, max_iter, sampling_method, verbose, defer_nonneg) -> bool:
    _check_fit_args.last_deferred = False
    deferred = False
    _require(isinstance(adata, AnnData), TypeError, "adata must be an AnnData object.")
    if is_sparse(adata.X):
        # extension over the reference (which raises the TypeError below for sparse input, main.py:395-396): a
        # scipy.sparse matrix takes the CSR tile-list path of the kernels (BASELINE north_star, config 4)
        _require(adata.X.ndim == 2, ValueError, "adata.X must be a 2D numpy array.")
        _require(_all_non_negative(np.asarray(adata.X.data)), ValueError, NONNEG_MSG)
    else:
        _require(isinstance(adata.X, np.ndarray), TypeError, "adata.X must be a numpy array.")
        _require(adata.X.ndim == 2, ValueError, "adata.X must be a 2D numpy array.")
        if defer_nonneg and adata.X.size > (1 << 24):
            deferred = _check_fit_args.last_deferred = True
        else:
            _require(_all_non_negative(adata.X), ValueError, NONNEG_MSG)
    _require(isinstance(covariate_keys, list), TypeError, "covariate_keys must be a list.")
    _require(len(covariate_keys) == len(m.n_covariate_components), ValueError,
             "Length of covariate_keys must match length of n_covariate_components.")
    for key in covariate_keys:
        _require(isinstance(key, str), TypeError, "Each element in covariate_keys must be a string.")
        _require(key in adata.obs.columns, ValueError, f"Covariate key '{key}' not found in adata.obs.")
        _require(adata.obs[key].dtype.kind == "O", TypeError,
                 f"Covariate '{key}' in adata.obs must be a categorical or object type variable.")
    _require(isinstance(sampling_method, str), TypeError, "sampling_method must be a string.")
    _require(isinstance(verbose, bool), TypeError, "verbose must be a boolean.")
    return deferred


_check_fit_args.last_deferred = False


# limits of the native library (csrc/mu_small_kernels.cuh kMaxCov, csrc/mu_gemm_sm100.cuh kMaxK): reported here, by
# name, before anything is uploaded -- the reference itself has no such limits
MAX_TOTAL_COMPONENTS = 256
MAX_TOTAL_COMPONENTS_ALS = 128
MAX_COVARIATES = 8


def check_native_limits(m) -> None:
    """Inputs the reference accepts but the B200 library cannot run; raised by ``fit`` with an explicit message."""
    _require(len(m.n_covariate_components) <= MAX_COVARIATES, ValueError,
             f"alpine_b200 supports at most {MAX_COVARIATES} covariates (got {len(m.n_covariate_components)}).")
    _require(all(k > 0 for k in m.n_covariate_components), ValueError,
             "alpine_b200 needs at least one component per covariate block (an entry of n_covariate_components is 0).")
    total = sum(m.n_covariate_components) + m.n_components
    _require(total <= MAX_TOTAL_COMPONENTS, ValueError,
             f"alpine_b200 supports at most {MAX_TOTAL_COMPONENTS} components in total (got {total}).")
    _require(not (m.use_als and total > MAX_TOTAL_COMPONENTS_ALS), ValueError,
             f"alpine_b200 supports at most {MAX_TOTAL_COMPONENTS_ALS} components in total with use_als=True (got {total}).")


def check_trained(m) -> None:
    _require(hasattr(m, "matrices"), RuntimeError, "Model is not trained yet. Please fit the model first.")


def check_adata(adata) -> None:
    _require(isinstance(adata, AnnData), TypeError, "adata must be an AnnData object.")


def check_transform_args(m, adata, n_iter) -> None:
    """main.py:154-164."""
    check_trained(m)
    check_adata(adata)
    ok = isinstance(n_iter, (int, type(None))) and not (n_iter is not None and n_iter <= 0)
    _require(ok, ValueError, "n_iter must be a positive integer or None.")


def check_range(rng: Any, name: str, lo_ok: Callable[[float], bool] = lambda v: True) -> Tuple[float, float]:
    """(low, high) tuple checks shared by the ComponentOptimizer validators (optimization.py:552-604)."""
    _require(isinstance(rng, (tuple, list)) and len(rng) == 2, ValueError, f"{name} must be a tuple of two values.")
    lo, hi = rng
    _require(lo <= hi, ValueError, f"{name} lower bound must be less than or equal to upper bound.")
    _require(lo_ok(lo), ValueError, f"{name} has an invalid lower bound.")
    return lo, hi


def as_key_list(keys: Sequence[str]) -> list:
    return list(keys)
