"""Drop-in ``ALPINE`` model class over the B200-native MU loop.

Public surface = the reference's ``alpine/main.py`` (class ``ALPINE`` main.py:46-320, ``AlpineMatrices``
main.py:28-43): same constructor keywords, methods, attributes set by ``fit`` and AnnData slots written.  The
arithmetic of the loop (``_fit`` main.py:486-676, ``_compute_loss`` 726-753, ``_scale_matrices`` 772-781, the
``_transform`` loop 705-709) runs in ``libalpine_b200.so`` through ``alpine_b200._native``; host code here only
prepares tensors (``_initialize_matrices`` keeps the reference's seeding and draw order, main.py:436-472, so the
initial W/H/B are bit-identical to the reference on the same device type) and moves results back.

There is no CPU path: ``device`` must be a CUDA device with an sm_100 GPU behind it.
"""
from __future__ import annotations

import gc
import os
import warnings
from contextlib import nullcontext
from copy import copy, deepcopy
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Union

import numpy as np
import numpy.typing as npt
import pandas as pd
import torch

from . import _native, validation
from .engine import MUEngine, dist_info, shard_bounds
from .utils.anndata_compat import AnnData, is_sparse
from .utils.encoder import FeatureEncoders
from .utils.kneedle import find_elbow

Float32Array = npt.NDArray[np.float32]


@dataclass
class AlpineMatrices:
    """Device-resident factor matrices (reference main.py:28-43).

    ``X`` is the genes x cells *view* of the cells-major upload (same strides as the reference's tensor,
    SURVEY.md 8 a1); ``Ws`` / ``Hs`` are column / row block views of the packed ``W`` (G x K) and ``H`` (K x n)
    buffers the kernels update in place; blocks are ordered covariates first, unguided last (main.py:79).
    Under cell sharding ``X``, ``Ys`` and ``Hs`` hold this rank's column block ``[lo, hi)``.
    """

    X: torch.Tensor
    Ys: List[torch.Tensor]
    Ws: List[torch.Tensor]
    Hs: List[torch.Tensor]
    Bs: List[torch.Tensor]
    W: Optional[torch.Tensor] = field(default=None, repr=False)
    H: Optional[torch.Tensor] = field(default=None, repr=False)
    X_cells_major: Optional[torch.Tensor] = field(default=None, repr=False)
    X_csr: Optional[tuple] = field(default=None, repr=False)  # sparse input: device (indptr, indices, values) over cells
    X_host: Optional[np.ndarray] = field(default=None, repr=False)
    Ys_host: Optional[List[np.ndarray]] = field(default=None, repr=False)
    shard: tuple = (0, 0)
    n_total: int = 0
    solver: Optional[object] = field(default=None, repr=False)  # native context kept between _fit and _scale_matrices

    def to_numpy(self, bufs: Optional["_HostBuffers"] = None) -> Dict[str, Union[Float32Array, List[Float32Array]]]:
        # the reference copies X back from the device (main.py:38); the host copy it came from is identical
        if self.X_host is not None:
            X = self.X_host
        elif self.X is not None:
            X = self.X.cpu().numpy().astype(np.float32)
        else:
            raise RuntimeError("sparse AlpineMatrices without a host copy of X")
        Ys = self.Ys_host if self.Ys_host is not None else [y.cpu().numpy().astype(np.float32) for y in self.Ys]
        if self.W is not None and self.H is not None:
            # one device -> host copy per packed factor (and one gather of the cell blocks under sharding); the blocks
            # are cut out on the host
            take = (lambda name, shape: bufs.take(name, shape)) if bufs is not None else (lambda name, shape: None)
            W = _to_host(self.W, take("W", tuple(self.W.shape)))
            Hd = _gather_cells([self.H], self.shard, self.n_total)[0]
            H = _to_host(Hd, take("H", tuple(Hd.shape)))
            cuts_w = np.cumsum([0] + [w.shape[1] for w in self.Ws])
            cuts_h = np.cumsum([0] + [h.shape[0] for h in self.Hs])
            Ws = []
            for i in range(len(self.Ws)):
                blk = W[:, cuts_w[i]:cuts_w[i + 1]]
                dst = take(f"W{i}", blk.shape)
                if dst is None:
                    Ws.append(np.ascontiguousarray(blk))
                else:
                    np.copyto(dst, blk)
                    Ws.append(dst)
            Hs = [H[cuts_h[i]:cuts_h[i + 1]] for i in range(len(self.Hs))]
        else:
            Ws = [w.cpu().numpy().astype(np.float32, copy=False) for w in self.Ws]
            Hs = [h.cpu().numpy().astype(np.float32, copy=False)
                  for h in _gather_cells(self.Hs, self.shard, self.n_total)]
        return {"X": X, "Ys": Ys, "Ws": Ws, "Hs": Hs,
                "Bs": [b.cpu().numpy().astype(np.float32, copy=False) for b in self.Bs]}


def _as_csr_f32(X):
    """Canonical fp32 CSR over cells of a scipy.sparse matrix (duplicates summed, indices sorted)."""
    import scipy.sparse as sp

    X = sp.csr_matrix(X, dtype=np.float32)
    X.sum_duplicates()
    return X


def _upload_csr(Xcsr, lo: int, hi: int, dev):
    """Rows [lo, hi) of a host CSR matrix as device (indptr int64 rebased to 0, indices int32, values fp32)."""
    p0, p1 = int(Xcsr.indptr[lo]), int(Xcsr.indptr[hi])
    indptr = torch.from_numpy(np.asarray(Xcsr.indptr[lo:hi + 1], dtype=np.int64) - p0).to(dev)
    indices = torch.from_numpy(np.ascontiguousarray(Xcsr.indices[p0:p1], dtype=np.int32)).to(dev)
    values = torch.from_numpy(np.ascontiguousarray(Xcsr.data[p0:p1], dtype=np.float32)).to(dev)
    return indptr, indices, values


def _csr_take_rows(csr, idx: torch.Tensor):
    """Rows ``idx`` (with repetition, in order) of a device CSR triple -> a new device CSR triple."""
    indptr, indices, values = csr
    starts = indptr[idx]
    counts = indptr[idx + 1] - starts
    new_indptr = torch.zeros(idx.numel() + 1, dtype=torch.int64, device=indptr.device)
    torch.cumsum(counts, 0, out=new_indptr[1:])
    total = int(new_indptr[-1].item())
    # position p of the output belongs to row r(p) = searchsorted(new_indptr, p) and is element p - new_indptr[r] of it
    row_of = torch.repeat_interleave(torch.arange(idx.numel(), device=indptr.device), counts, output_size=total)
    src = starts[row_of] + (torch.arange(total, device=indptr.device) - new_indptr[row_of])
    return new_indptr, indices[src].contiguous(), values[src].contiguous()


class _Background:
    """A callable running on its own thread (the native calls inside release the GIL); ``result()`` joins and
    re-raises."""

    def __init__(self, fn):
        import threading

        self._out, self._exc = None, None
        self._thread = threading.Thread(target=self._run, args=(fn,), daemon=True)
        self._thread.start()

    def _run(self, fn):
        try:
            self._out = fn()
        except BaseException as exc:  # re-raised by result()
            self._exc = exc

    def result(self):
        self._thread.join()
        if self._exc is not None:
            raise self._exc
        return self._out


class _HostBuffers:
    """Host arrays for the results of a fit, allocated and touched on a background thread while the upload and the
    loop run: a fresh 40 MB numpy array costs about as much in page faults as in the copy that fills it, and the
    faults would otherwise sit between the last iteration and the return of ``fit``.  ``take`` hands a buffer out once
    (``None`` if the shape was not prepared); the arrays are ordinary numpy arrays owned by whoever takes them."""

    def __init__(self, specs: Dict[str, tuple]):
        import ctypes

        def work():
            out = {}
            for name, shape in specs.items():
                a = np.empty(shape, dtype=np.float32)
                if a.nbytes:
                    ctypes.memset(a.ctypes.data, 0, a.nbytes)  # touches every page; ctypes releases the GIL
                out[name] = a
            return out

        self._job = _Background(work)
        self._bufs: Optional[dict] = None

    def take(self, name: str, shape: tuple) -> Optional[np.ndarray]:
        if self._bufs is None:
            self._bufs = self._job.result()
        a = self._bufs.pop(name, None)
        return a if a is not None and a.shape == tuple(shape) else None


def _to_host(t: torch.Tensor, out: Optional[np.ndarray]) -> np.ndarray:
    """Device tensor -> numpy array, into ``out`` when a prepared buffer of the right shape is at hand."""
    if out is None:
        return t.cpu().numpy()
    torch.from_numpy(out).copy_(t)
    return out


def _gather_cells(Hs: List[torch.Tensor], shard, n_total: int) -> List[torch.Tensor]:
    """All ranks' column blocks of every H block, concatenated along cells (identity without sharding)."""
    rank, world = dist_info()
    if world == 1 or n_total == 0 or (shard[1] - shard[0]) == n_total:
        return Hs
    import torch.distributed as dist

    out = []
    for h in Hs:
        full = torch.zeros((h.shape[0], n_total), dtype=h.dtype, device=h.device)
        full[:, shard[0]:shard[1]] = h
        dist.all_reduce(full)  # disjoint blocks: the sum is the concatenation
        out.append(full)
    return out


class ALPINE:
    def __init__(
        self,
        n_components: int,
        n_covariate_components: List[int],
        lam: List[float],
        orth_W: float = 0.0,
        alpha_W: float = 0.0,
        l1_ratio_W: float = 0.0,
        use_als: bool = False,
        scale_needed: bool = True,
        loss_type: str = "kl-divergence",
        device: str = "cuda",
        eps: float = 1e-6,
        random_state: int = 42,
        l1_ratio: Optional[float] = None,
    ):
        # `l1_ratio` is accepted as an alias of the reference's `l1_ratio_W` (BASELINE north_star spelling)
        if l1_ratio is not None:
            l1_ratio_W = l1_ratio
        self.n_components = n_components
        self.n_covariate_components = n_covariate_components
        self.lam = lam
        self.orth_W = orth_W
        self.alpha_W = alpha_W
        self.l1_ratio_W = l1_ratio_W
        self.use_als = use_als
        self.scale_needed = scale_needed
        self.device = torch.device(device)
        self.loss_type = loss_type
        self.eps = eps
        self.random_state = random_state
        validation.check_model_args(self)
        self.n_all_components = self.n_covariate_components + [self.n_components]  # main.py:79
        self.total_components = sum(self.n_all_components)

    # ------------------------------------------------------------------------------------------------ fit
    def fit(
        self,
        adata: AnnData,
        covariate_keys: List[str],
        batch_size: Optional[int] = None,
        max_iter: Optional[int] = None,
        sampling_method: str = "random",
        verbose: bool = False,
    ) -> "ALPINE":
        import time

        t0 = time.perf_counter()
        self.timings: Dict[str, float] = {}

        def lap(name: str) -> None:
            nonlocal t0
            if self.device.type == "cuda" and torch.cuda.is_available():
                torch.cuda.synchronize(self._cuda_device())
            t1 = time.perf_counter()
            self.timings[name] = self.timings.get(name, 0.0) + (t1 - t0)
            t0 = t1

        # the non-negativity scan of a large dense X is done on the device after the upload (same ValueError)
        self._nonneg_pending = validation.check_fit_args(self, adata, covariate_keys, batch_size, max_iter,
                                                         sampling_method, verbose, defer_nonneg=True)
        validation.check_native_limits(self)
        lap("validate")
        self.feature_names = adata.var_names.tolist()
        self.n_features = adata.shape[1]
        self.covariate_keys = covariate_keys
        self.sampling_method = sampling_method
        self.verbose = verbose

        # The reference transposes to genes x cells (main.py:104); that view shares the cells-major buffer, which
        # is the layout the kernels stream, so the "transpose" stays a view here as well.
        if is_sparse(adata.X):
            # CSR over cells (AnnData's layout for count matrices); .T is the genes x cells view of the same data
            X = _as_csr_f32(adata.X).T
        else:
            X = np.ascontiguousarray(adata.X, dtype=np.float32).T
        n_sample = X.shape[1]
        # result arrays (W, H, their per-block copies for the AnnData slots) are allocated and touched in the background
        ks = list(self.n_all_components)
        specs = {"W": (self.n_features, sum(ks)), "H": (sum(ks), n_sample)}
        for i, k in enumerate(ks):
            specs[f"W{i}"] = (self.n_features, k)
            specs[f"varm{i}"] = (self.n_features, k)
            specs[f"obsm{i}"] = (k, n_sample)
        bufs = _HostBuffers(specs)
        # the upload of this rank's cells starts now, on host threads of the library, and overlaps the label encoding
        # and the factor draws; a warm-up fit and the main fit share the one device copy of X (it is never written)
        Xdev = self._start_upload(X)
        self.fe = FeatureEncoders(covariate_keys)
        Y = self.fe.fit_transform(adata.obs)
        self.batch_size = batch_size if batch_size is not None else n_sample
        lap("encode")

        if max_iter is None:
            # warm-up run + Kneedle elbow on log10(reconstruction loss) (main.py:116-131, 755-770)
            m_warmup = self._initialize_matrices(X, Y, _Xdev=Xdev)
            self.max_iter = 200
            self._fit(m_warmup)
            self.max_iter = self._compute_best_iter(self.loss_history["reconstruction loss"].values)
            del m_warmup
            gc.collect()
            torch.cuda.empty_cache()
            lap("warmup_fit")
        else:
            self.max_iter = max_iter

        m = self._initialize_matrices(X, Y, _Xdev=Xdev)
        del Xdev
        lap("upload_init")
        try:
            self._fit(m, keep_solver=True)
            lap("loop")
            if self.scale_needed:
                self._scale_matrices(m)
        finally:
            if m.solver is not None:
                m.solver.close()
                m.solver = None
        self.matrices = m.to_numpy(bufs)
        self._rng_publish()
        lap("scale_download")
        self.store_embeddings(adata, _dummy_matrices=Y, _bufs=bufs)  # the encoders were fitted on this adata a moment ago
        lap("store_embeddings")
        return self

    def fit_transform(
        self,
        adata: AnnData,
        covariate_keys: List[str],
        batch_size: Optional[int] = None,
        max_iter: Optional[int] = None,
        sampling_method: str = "random",
        verbose: bool = False,
    ) -> None:
        self.fit(adata, covariate_keys, batch_size=batch_size, max_iter=max_iter, sampling_method=sampling_method,
                 verbose=verbose).transform(adata)

    # ------------------------------------------------------------------------------------------ transform
    def transform(self, adata: AnnData, n_iter: Optional[int] = None) -> None:
        validation.check_transform_args(self, adata, n_iter)
        self._transform(adata, n_iter if n_iter is not None else self.max_iter)

    def _transform(self, adata: AnnData, n_iter: int) -> None:
        """H-only MU on new cells with the fitted (scaled) W (main.py:678-724).

        The reference recomputes the loop-invariant ``2 W^T X`` every iteration; here ``A = W^T X`` (one sweep of X)
        and ``T = W^T W`` are formed once and each iteration is ``H *= 2A / max(2 T H, eps)`` on K x n only.
        """
        sparse = is_sparse(adata.X)
        Xcm = _as_csr_f32(adata.X) if sparse else np.ascontiguousarray(adata.X, dtype=np.float32)
        if not np.all((Xcm.data if sparse else Xcm) >= 0):
            raise ValueError("All elements in adata.X must be non-negative.")
        dev = self._cuda_device()
        n_sample, G = Xcm.shape
        K = self.total_components
        if not sparse:
            Xd = _native.padded_rows(n_sample, G, dev)
            _native.upload_rows(Xd, Xcm)
        # un-reseeded draw from the device generator, as the reference (main.py:687-689): the process-wide one, which
        # fit() left where the reference leaves it -- or this model's own stream when it was told to keep it private
        H = _native.padded_rows(K, n_sample, dev)
        gen = getattr(self, "_gen_dev", None) if getattr(self, "_rng_private", False) else None
        if gen is not None and gen.device != dev:
            gen = None
        H.copy_(torch.rand((K, n_sample), dtype=torch.float32, device=dev, generator=gen))
        W = torch.cat([torch.tensor(w, dtype=torch.float32, device=dev) for w in self.matrices["Ws"]], dim=1).contiguous()
        solver = _native.Solver(dev, G, n_sample, [K], [])
        try:
            if sparse:
                solver.bind_csr(*_upload_csr(Xcm, 0, n_sample, dev))
            else:
                solver.bind_dense(Xd)
            solver.bind_factors(W, H, [])
            solver.set_hparams([], 0.0, 0.0, 0.0, self.eps)
            solver.transform(n_iter)
            H_host = H.cpu().numpy()
        finally:
            solver.close()
        start = 0
        parts = []
        for k in self.n_all_components:
            parts.append(H_host[start:start + k])
            start += k
        for i, covariate in enumerate(self.covariate_keys):
            adata.obsm[covariate] = parts[i].T
            adata.varm[covariate] = deepcopy(self.matrices["Ws"][i])
        adata.obsm["ALPINE_embedding"] = parts[-1].T
        adata.varm["ALPINE_weights"] = deepcopy(self.matrices["Ws"][-1])

    # -------------------------------------------------------------------------------------------- queries
    def compute_loss(self, adata: AnnData):
        """Host re-evaluation of the objective on (possibly new) data (main.py:187-236)."""
        validation.check_trained(self)
        validation.check_adata(adata)
        if "ALPINE_embedding" not in adata.obsm:
            raise ValueError("ALPINE_embedding not found in adata.obsm. Please transform the data first.")
        X = (adata.X.toarray() if is_sparse(adata.X) else np.asarray(adata.X)).astype(np.float32).T
        Hs = [np.asarray(adata.obsm[c]).T for c in self.covariate_keys] + [np.asarray(adata.obsm["ALPINE_embedding"]).T]
        Ws = [np.asarray(adata.varm[c]) for c in self.covariate_keys] + [np.asarray(adata.varm["ALPINE_weights"])]
        W, H = np.concatenate(Ws, axis=1), np.concatenate(Hs, axis=0)
        recon_loss = np.linalg.norm(X - W @ H, ord="fro") ** 2
        Ys = self.fe.transform(adata.obs)
        Bs = self.matrices["Bs"]
        pred_loss = []
        for i in range(len(Ys)):
            y, y_hat = Ys[i].T, Bs[i] @ Hs[i]
            if self.loss_type == "kl-divergence":
                y_hat = np.clip(y_hat, a_min=self.eps, a_max=None)
                pred_loss.append(np.sum(y * np.log(np.clip(y / y_hat, a_min=self.eps, a_max=None)) - y + y_hat))
            else:
                pred_loss.append(np.linalg.norm(y - y_hat, ord="fro") ** 2)
        return recon_loss + sum(self.lam[i] * pl for i, pl in enumerate(pred_loss))

    def get_decomposed_matrices(self):
        validation.check_trained(self)
        return self.matrices

    def get_covariate_gene_scores(self, adata: Optional[AnnData] = None) -> Union[Dict[str, pd.DataFrame], None]:
        """Per-category mean embedding pushed through W_i (main.py:246-273)."""
        validation.check_trained(self)
        scores = {}
        for i, covariate in enumerate(self.covariate_keys):
            W, H, Y = self.matrices["Ws"][i], self.matrices["Hs"][i], self.matrices["Ys"][i]
            per_category = H @ Y.T / Y.sum(axis=1)
            scores[covariate] = pd.DataFrame(W @ per_category, index=self.feature_names,
                                             columns=self.fe.encoded_labels[covariate])
        if adata is None:
            return scores
        for condition, df in scores.items():
            adata.varm[condition + "_gene_scores"] = df
        return None

    def get_normalized_expression(self, adata: AnnData, library_size: Optional[float] = None) -> None:
        """Reconstruct counts from the unguided block and library-size normalise them (main.py:275-301).

        ``scanpy.pp.normalize_total`` (absent here) is restated: every cell is scaled to ``target_sum``, which
        defaults to the median of the per-cell totals; cells with a zero total are left untouched.
        """
        validation.check_trained(self)
        validation.check_adata(adata)
        if "ALPINE_embedding" not in adata.obsm:
            raise ValueError("ALPINE_embedding not found in adata.obsm. Please transform the data first.")
        if (library_size is not None) and (library_size <= 0):
            raise ValueError("library_size must be a positive float.")
        W = self.matrices["Ws"][-1]
        H = np.asarray(adata.obsm["ALPINE_embedding"]).T
        Xn = np.dot(W, H).astype(np.float32).T
        totals = Xn.sum(axis=1)
        target = float(np.median(totals[totals > 0])) if library_size is None else float(library_size)
        scale = np.where(totals > 0, totals / target, 1.0).astype(np.float32)
        adata.layers["normalized_expression"] = Xn / scale[:, None]

    def store_embeddings(self, adata: AnnData, _dummy_matrices=None, _bufs=None) -> None:
        """Write embeddings / weights into the AnnData slots of main.py:303-320 (copies, as in the reference)."""
        validation.check_trained(self)
        validation.check_adata(adata)

        def dup(a: np.ndarray, name: str, transposed: bool) -> np.ndarray:
            # copy(a.T) keeps a's memory order: a (k, n) C-ordered block becomes an (n, k) F-ordered array
            dst = _bufs.take(name, a.shape) if _bufs is not None else None
            if dst is None:
                return copy(a.T) if transposed else copy(a)
            np.copyto(dst, a)
            return dst.T if transposed else dst

        last = len(self.matrices["Hs"]) - 1
        adata.obsm["ALPINE_embedding"] = dup(self.matrices["Hs"][-1], f"obsm{last}", True)
        adata.varm["ALPINE_weights"] = dup(self.matrices["Ws"][-1], f"varm{last}", False)
        dummy_matrices = _dummy_matrices if _dummy_matrices is not None else self.fe.transform(adata.obs)
        for i, covariate in enumerate(self.covariate_keys):
            adata.obsm[covariate] = dup(self.matrices["Hs"][i], f"obsm{i}", True)
            adata.obsm[f"{covariate}_dummy_matrix"] = dummy_matrices[i]
            adata.varm[covariate] = dup(self.matrices["Ws"][i], f"varm{i}", False)

    # ------------------------------------------------------------------------------------- device helpers
    def _cuda_device(self) -> torch.device:
        if self.device.type != "cuda":
            raise _native.AlpineNativeError(
                f"alpine_b200 runs the MU loop on a B200 only (device={str(self.device)!r}); there is no CPU fallback")
        if not torch.cuda.is_available():
            raise _native.AlpineNativeError("no CUDA device is available; alpine_b200 has no CPU fallback")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        return torch.device("cuda", idx)

    def _rng_begin(self, dev: torch.device) -> torch.Generator:
        """Fresh per-model generators with the seeds main.py:440-442 gives the process-wide ones."""
        self._gen_dev = torch.Generator(device=dev)
        self._gen_dev.manual_seed(self.random_state)
        self._gen_cpu = torch.Generator()
        self._gen_cpu.manual_seed(self.random_state)
        return self._gen_dev

    def _rng_skip_randperms(self, n: int, count: int) -> None:
        """Advance the device generator as ``count`` calls of ``torch.randperm(n, device=...)`` would (the reference
        draws one permutation per full-batch iteration, main.py:502-506, which only reorders fp32 sums there and is
        not needed here).  One real draw measures the Philox offset a call consumes; the rest is arithmetic."""
        g = getattr(self, "_gen_dev", None)
        if g is None or count <= 0:
            return
        try:
            o0 = g.get_offset()
            torch.randperm(n, device=g.device, generator=g)
            o1 = g.get_offset()
            g.set_offset(o1 + (count - 1) * (o1 - o0))
        except (AttributeError, RuntimeError):  # a torch without offset access: draw them all
            for _ in range(count - 1):
                torch.randperm(n, device=g.device, generator=g)

    def _rng_publish(self) -> None:
        """Leave the process-wide generators where the reference's ``fit`` leaves them (unless this model was told to
        keep its streams private: ``_rng_private``, set by ComponentOptimizer's worker threads)."""
        if getattr(self, "_rng_private", False) or getattr(self, "_gen_dev", None) is None:
            return
        idx = self._gen_dev.device.index
        torch.cuda.default_generators[idx if idx is not None else torch.cuda.current_device()].set_state(
            self._gen_dev.get_state())
        torch.default_generator.set_state(self._gen_cpu.get_state())

    def _start_upload(self, X_array):
        """Begin the host -> device copy of this rank's block of cells on a background thread; the returned handle's
        ``result()`` is what ``_initialize_matrices`` takes as ``_Xdev``: the cells-major device matrix, or the device
        CSR triple for sparse input."""
        dev = self._cuda_device()
        G, n = X_array.shape
        rank, world = dist_info()
        if world > n:
            raise ValueError(f"cell sharding needs at least one cell per rank ({n} cells, {world} ranks)")
        lo, hi = shard_bounds(n, world, rank)
        Xcm_host = X_array.T  # cells x genes; C-contiguous when X_array came from fit()

        detail = self.__dict__.setdefault("timings_detail", {})

        def work():
            import time

            t0 = time.perf_counter()
            if is_sparse(X_array):
                return _upload_csr(_as_csr_f32(Xcm_host), lo, hi, dev)
            Xd = _native.padded_rows(hi - lo, G, dev)
            t1 = time.perf_counter()
            _native.upload_rows(Xd, Xcm_host[lo:hi])
            detail["x_alloc"], detail["x_upload_call"] = t1 - t0, time.perf_counter() - t1
            return Xd

        return _Background(work)

    def _initialize_matrices(self, X_array: Float32Array, Y_list_array: List[Float32Array],
                             _Xdev=None) -> AlpineMatrices:
        """Seed, upload, draw W / H / B in the reference's order (main.py:436-472).

        ``X_array`` is genes x cells (a view of the cells-major buffer).  With torch.distributed initialised and
        world_size > 1, every rank draws the full W, H, B from the same seed (identical streams on identical
        devices) and keeps its own column block of X, Y and H.
        """
        dev = self._cuda_device()
        # The reference seeds the process-wide generators here (torch.manual_seed / torch.cuda.manual_seed) and every
        # later draw -- W, H, B below, one randperm per iteration, the weighted sampler, transform's H0 -- continues
        # those streams.  The same streams are kept here in two per-model generators (same seed, same draw order =>
        # same numbers), because ComponentOptimizer runs one fit per GPU from worker threads and the global seeding
        # calls would race between them; ``fit`` publishes their final state to the process-wide generators, so a
        # later ``transform`` (or any user draw) starts exactly where it starts after the reference's ``fit``.
        gen = self._rng_begin(dev)
        G, n = X_array.shape
        rank, world = dist_info()
        if world > n:
            raise ValueError(f"cell sharding needs at least one cell per rank ({n} cells, {world} ranks)")
        lo, hi = shard_bounds(n, world, rank)
        n_loc = hi - lo
        K = self.total_components

        import time

        detail = self.__dict__.setdefault("timings_detail", {})
        t_begin = time.perf_counter()
        if _Xdev is None:
            _Xdev = self._start_upload(X_array)
        Ys_host = [np.ascontiguousarray(y.T, dtype=np.float32) for y in Y_list_array]  # c_i x n (main.py:447)
        Ys = [torch.from_numpy(np.ascontiguousarray(y[:, lo:hi])).to(dev) for y in Ys_host]

        W = torch.empty((G, K), dtype=torch.float32, device=dev)
        H = _native.padded_rows(K, n_loc, dev)
        eps = self.eps
        col = 0
        Ws = []
        for k in self.n_all_components:  # main.py:454-458
            W[:, col:col + k] = torch.rand((G, k), dtype=torch.float32, device=dev, generator=gen).clamp(min=eps)
            Ws.append(W[:, col:col + k])
            col += k
        row = 0
        Hs = []
        for k in self.n_all_components:  # main.py:460-464
            full = torch.rand((k, n), dtype=torch.float32, device=dev, generator=gen).clamp(min=eps)
            H[row:row + k, :] = full[:, lo:hi]
            Hs.append(H[row:row + k, :])
            row += k
            del full
        Bs = [torch.rand((y.shape[0], k), dtype=torch.float32, device=dev, generator=gen).clamp(min=eps).contiguous()
              for (y, k) in zip(Ys_host, self.n_covariate_components)]  # main.py:466-470

        # X: uploaded by now, or still arriving while the draws above ran
        t_draws = time.perf_counter()
        uploaded = _Xdev.result() if isinstance(_Xdev, _Background) else _Xdev
        detail["init_labels_draws"], detail["init_wait_upload"] = t_draws - t_begin, time.perf_counter() - t_draws
        X_csr, Xd = (uploaded, None) if isinstance(uploaded, tuple) else (None, uploaded)
        if Xd is not None and getattr(self, "_nonneg_pending", False):
            self._nonneg_pending = False
            ok = torch.tensor([1.0 if (n_loc == 0 or bool(Xd.min() >= 0)) else 0.0], device=dev)  # NaN fails
            if world > 1:
                import torch.distributed as dist

                dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # every rank raises, or none
            if float(ok.item()) != 1.0:
                raise ValueError(validation.NONNEG_MSG)
        return AlpineMatrices(X=Xd.T if Xd is not None else None, Ys=Ys, Ws=Ws, Hs=Hs, Bs=Bs, W=W, H=H,
                              X_cells_major=Xd, X_csr=X_csr, X_host=X_array, Ys_host=Ys_host, shard=(lo, hi), n_total=n)

    def _make_solver(self, m: AlpineMatrices) -> "_native.Solver":
        n_loc, G = m.H.shape[1], m.W.shape[0]
        solver = _native.Solver(m.W.device, G, n_loc, self.n_all_components, [y.shape[0] for y in m.Ys], self.loss_type)
        if m.X_csr is not None:
            solver.bind_csr(*m.X_csr)
        else:
            solver.bind_dense(m.X_cells_major)
        solver.bind_labels(m.Ys)
        solver.bind_factors(m.W, m.H, m.Bs)
        solver.set_hparams(self.lam, self.alpha_W, self.l1_ratio_W, self.orth_W, self.eps)
        return solver

    # ------------------------------------------------------------------------------------------- hot loop
    def _fit(self, m: AlpineMatrices, keep_solver: bool = False) -> None:
        """``max_iter`` MU iterations, in place on ``m`` (main.py:486-676); leaves ``self.loss_history``.

        ``keep_solver`` leaves the native context (workspaces, plans) on ``m.solver`` for ``_scale_matrices``."""
        full_batch = self.batch_size >= m.n_total
        if self.sampling_method not in ("random", "weighted"):
            raise ValueError(f"Unknown sampling method: {self.sampling_method}. Only 'weighted', and 'random' are supported.")
        colnames = ["total loss", "reconstruction loss"] + [f"prediction loss({k})" for k in self.covariate_keys]
        if self.max_iter <= 0:  # the reference's loop body never runs: empty loss history, factors untouched
            self.loss_history = pd.DataFrame([], columns=colnames)
            return
        if not full_batch or self.sampling_method == "weighted":
            return self._fit_minibatch(m)
        import time

        detail = self.__dict__.setdefault("timings_detail", {})
        t0 = time.perf_counter()
        self._rng_skip_randperms(m.n_total, self.max_iter)  # main.py:502-506, one permutation per iteration
        solver = self._make_solver(m)
        detail["loop_make_solver"] = time.perf_counter() - t0
        try:
            if dist_info()[1] > 1 and not self.use_als and os.environ.get("ALPINE_B200_PEER", "0") == "1":
                # Opt-in: exchange over NVLink peer memory inside the W-update kernels instead of the NCCL
                # all-reduce (falls back to it, on every rank, when CUDA IPC is not available).  Mapping the peers'
                # exchange blocks costs ~0.2 s per fit and saves ~35 us per iteration at 8 GPUs, so it only pays
                # for very long fits; bench.py's device-resident arm uses it.
                solver.enable_peer_exchange()
            engine = MUEngine(solver, self.lam, use_als=self.use_als)
            pbar = None
            if self.verbose:
                from tqdm import tqdm

                pbar = tqdm(total=self.max_iter, desc="Iteration", ncols=100)
            with (pbar if pbar is not None else nullcontext()):
                t1 = time.perf_counter()
                engine.begin(self.max_iter)
                t2 = time.perf_counter()
                for it in range(self.max_iter):
                    engine.step(it)
                    if pbar is not None:
                        pbar.update(1)
                t3 = time.perf_counter()
                history = engine.collect_losses(self.max_iter)
                detail["loop_begin"], detail["loop_enqueue"], detail["loop_collect"] = t2 - t1, t3 - t2, time.perf_counter() - t3
        finally:
            if keep_solver:
                m.solver = solver
            else:
                solver.close()
        self.loss_history = pd.DataFrame(history.tolist(), columns=colnames)

    def _fit_minibatch(self, m: AlpineMatrices) -> None:
        """Epochs of mini-batch MU steps (main.py:500-521, 589-663) with the reference's index streams: the epoch
        indices come from the same generator calls in the same stream position as the reference's (``randperm`` on the
        device generator after the W / H / B draws; the weighted sampler's ``multinomial`` on the CPU generator).

        Per batch the cells ``idx`` are gathered into contiguous buffers (rows of the cells-major X -- or of the CSR
        matrix --, columns of H and Y: one launch, ``alpine_batch_gather``), one MU step runs on them with the same
        kernels as the full-batch path, and the H columns are scattered back (``Hs[j][:, idx] = ...``, main.py:662,
        ``alpine_batch_scatter``; duplicate indices of the weighted sampler carry identical columns).  The loss of every epoch is evaluated on the full data (main.py:666).
        Under cell sharding every rank draws the same epoch indices, takes the batch's cells that fall into its own
        column block and the batch's partial sums are all-reduced exactly like a full-batch iteration's.
        """
        from .utils.sampling import (create_joint_labels_from_dummy_matrices, generate_epoch_indices,
                                     get_batch_indices, get_num_batches)

        rank, world = dist_info()
        sharded = world > 1
        lo, hi = m.shard if sharded else (0, m.H.shape[1])
        n_total = m.n_total if sharded else m.H.shape[1]
        dev = m.W.device
        n, G = m.H.shape[1], m.W.shape[0]
        sparse = m.X_csr is not None
        K = self.total_components
        if not m.Ys:
            joint_labels = [""] * n_total
        elif sharded:  # the sampler works on all cells: every rank draws the same epoch indices from the same stream
            joint_labels = create_joint_labels_from_dummy_matrices([torch.from_numpy(y) for y in m.Ys_host])
        else:
            joint_labels = create_joint_labels_from_dummy_matrices(m.Ys)
        bs = int(min(self.batch_size, n_total))
        solvers: dict = {}

        def batch_solver(size: int):
            if size not in solvers:
                Xb = None if sparse else _native.padded_rows(size, G, dev)
                Hb = _native.padded_rows(K, size, dev)
                Yb = [torch.empty((y.shape[0], size), dtype=torch.float32, device=dev) for y in m.Ys]
                s = _native.Solver(dev, G, size, self.n_all_components, [y.shape[0] for y in m.Ys], self.loss_type)
                if not sparse:
                    s.bind_dense(Xb)
                s.bind_labels(Yb)
                s.bind_factors(m.W, Hb, m.Bs)
                s.set_hparams(self.lam, self.alpha_W, self.l1_ratio_W, self.orth_W, self.eps)
                if sharded:
                    s.reduce_buffer()  # the batch's [X H^T | H H^T | ...] partials are all-reduced over the ranks
                solvers[size] = (s, Xb, Hb, Yb)
            return solvers[size]

        def all_reduce(t):
            import torch.distributed as dist

            dist.all_reduce(t)

        history = []
        pbar = None
        if self.verbose:
            from tqdm import tqdm

            pbar = tqdm(total=self.max_iter, desc="Iteration", ncols=100)
        # tests replay the reference's recorded sampler output through this hook (one index vector per epoch)
        stream = getattr(self, "_epoch_index_stream", None)
        full = self._make_solver(m)  # full-data context for the per-epoch loss (main.py:666)
        owner = None  # the batch context whose W^T master copy is ahead of m.W (consecutive batches of one size stay in it)
        try:
            full.fit_begin(1)
            xnorm2 = full.losses(0)[0]
            for _ in range(self.max_iter):
                if stream is not None:
                    epoch_indices = torch.as_tensor(next(stream), dtype=torch.long, device=dev)
                else:
                    epoch_indices = generate_epoch_indices(joint_labels=joint_labels,
                                                           sampling_method=self.sampling_method, device=dev,
                                                           generator=getattr(self, "_gen_dev", None),
                                                           cpu_generator=getattr(self, "_gen_cpu", None))
                for b in range(get_num_batches(len(epoch_indices), bs)):
                    idx = get_batch_indices(epoch_indices, b, bs)
                    if len(idx) == 0:
                        break
                    if sharded:
                        # this rank's cells of the batch, padded with empty cells (zero rows of X, zero columns of H
                        # and Y contribute nothing to any sum and stay zero) up to a multiple of 256, so that a handful
                        # of contexts serves every batch; the exchange buffers have the same size on every rank
                        loc = idx[(idx >= lo) & (idx < hi)] - lo
                        cnt = int(loc.numel())
                        size = max(256, (cnt + 255) // 256 * 256)
                    else:
                        loc, cnt, size = idx, len(idx), len(idx)
                    s, Xb, Hb, Yb = batch_solver(size)
                    if owner is not None and owner is not s:
                        owner.sync_w()  # another context takes over: it reads the shared row-major W
                        owner = None
                    loc = loc.contiguous()
                    if sparse:  # the batch's cells as their own CSR matrix -> tile lists (once per batch)
                        indptr, indices, values = _csr_take_rows(m.X_csr, loc)
                        if cnt < size:
                            indptr = torch.cat([indptr, indptr[-1:].expand(size - cnt)])
                        s.bind_csr(indptr.contiguous(), indices, values)
                    # one launch: rows of the cells-major X (contiguous), columns of H and of every Y, zero padding
                    s.batch_gather(None if sparse else m.X_cells_major, m.H, m.Ys, loc)
                    s.batch_begin()
                    s.mu_partials()
                    if sharded:
                        all_reduce(s.reduce_buffer())
                    if self.use_als:  # main.py:523-588 on the batch
                        for blk in range(s.n_blocks):
                            s.als_block(blk)
                            if sharded and blk + 1 < s.n_blocks:
                                all_reduce(s.gram_view())
                        s.als_finish(0)
                    else:
                        s.mu_apply(0)
                        owner = s  # its W^T is now ahead of the shared row-major W
                    s.batch_scatter(m.H, loc)  # main.py:662; duplicates of the weighted sampler carry identical columns
                if owner is not None:
                    owner.sync_w()  # the per-epoch loss (and the caller) read the shared row-major W
                    owner = None
                history.append(self._compute_loss(m, solver=full, xnorm2=xnorm2))
                if pbar is not None:
                    pbar.set_postfix({"objective loss": history[-1][0]})
                    pbar.update(1)
            for s, *_ in solvers.values():
                s.losses(0)  # synchronise and surface kernel faults
        finally:
            if pbar is not None:
                pbar.close()
            full.close()
            for s, *_ in solvers.values():
                s.close()
        colnames = ["total loss", "reconstruction loss"] + [f"prediction loss({k})" for k in self.covariate_keys]
        self.loss_history = pd.DataFrame(history, columns=colnames)

    def _compute_loss(self, m: AlpineMatrices, solver=None, xnorm2: Optional[float] = None) -> List[float]:
        """[total, reconstruction, prediction...] of the factors as they are (main.py:726-753).

        Not on the hot path of a full-batch fit (the loop gets its loss terms from the update kernels); mini-batch fits
        call it once per epoch.  Everything runs in the library (``alpine_eval_loss``): the reconstruction term by the
        same trace identity, ||X||^2 - 2 tr(W^T X H^T) + tr(W^T W H H^T), with W^T X, W^T W and H H^T from the tcgen05
        contraction and fp64 traces; the prediction terms by the statistics kernel.
        """
        own = solver is None
        if own:
            solver = self._make_solver(m)
        try:
            if xnorm2 is None:
                solver.fit_begin(1)  # ||X||^2 by the library's fp64 reduction (no fp64 copy of X)
                xnorm2 = solver.losses(0)[0]
            terms = solver.eval_loss()
        finally:
            if own:
                solver.close()
        vals = np.concatenate([[float(xnorm2)], terms])
        if dist_info()[1] > 1:
            import torch.distributed as dist

            t = torch.from_numpy(vals).to(m.W.device)
            dist.all_reduce(t)
            vals = t.cpu().numpy()
        vals = vals.tolist()
        recon = vals[0] - 2.0 * vals[1] + vals[2]
        preds = vals[3:]
        return [recon + sum(self.lam[i] * p for i, p in enumerate(preds)), recon] + preds

    def _compute_best_iter(self, train_loss) -> int:
        """Kneedle elbow of log10(reconstruction loss) (main.py:755-770)."""
        elbow = find_elbow(np.arange(0, len(train_loss)), np.log10(train_loss))
        if elbow is not None:
            return int(elbow)
        warnings.warn("Kneedle elbow not found, using default max_iter=200")
        return 200

    def _scale_matrices(self, m: AlpineMatrices) -> None:
        """Column-normalise every W block; H and B absorb the scale (main.py:772-781)."""
        own = m.solver is None
        solver = self._make_solver(m) if own else m.solver
        try:
            solver.scale()
            torch.cuda.synchronize(m.W.device)
        finally:
            if own:
                solver.close()
