"""ctypes binding of ``libalpine_b200.so`` (the C ABI declared in ``include/alpine_b200.h``).

PyTorch owns every tensor; this module only passes raw device pointers and the
current CUDA stream across the boundary.  There is no CPU fallback: if the
shared library is missing or no sm_100 GPU is present the calls raise.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import List, Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libalpine_b200.so")

LOSS_KL = 0
LOSS_FROBENIUS = 1
LOSS_TYPES = {"kl-divergence": LOSS_KL, "frobenius": LOSS_FROBENIUS}

_lib = None

# name -> (restype, argtypes); mirrors include/alpine_b200.h one to one
_c_ctx = ctypes.c_void_p
_f32p = ctypes.c_void_p
SIGNATURES = {
    "alpine_abi_version": (ctypes.c_int, []),
    "alpine_last_error": (ctypes.c_char_p, []),
    "alpine_launch_count": (ctypes.c_longlong, []),
    "alpine_create": (ctypes.c_int, [ctypes.POINTER(_c_ctx), ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                     ctypes.c_int]),
    "alpine_destroy": (ctypes.c_int, [_c_ctx]),
    "alpine_workspace_bytes": (ctypes.c_int64, [_c_ctx]),
    "alpine_bind_workspace": (ctypes.c_int, [_c_ctx, ctypes.c_void_p, ctypes.c_int64]),
    "alpine_bind_dense": (ctypes.c_int, [_c_ctx, _f32p, ctypes.c_int64]),
    "alpine_bind_labels": (ctypes.c_int, [_c_ctx, ctypes.c_int, _f32p]),
    "alpine_bind_factors": (ctypes.c_int, [_c_ctx, _f32p, ctypes.c_int64, _f32p, ctypes.c_int64,
                                           ctypes.POINTER(ctypes.c_void_p)]),
    "alpine_set_hparams": (ctypes.c_int, [_c_ctx, ctypes.POINTER(ctypes.c_double), ctypes.c_double, ctypes.c_double,
                                          ctypes.c_double, ctypes.c_double]),
    "alpine_reduce_buffer_size": (ctypes.c_int64, [_c_ctx]),
    "alpine_bind_reduce_buffer": (ctypes.c_int, [_c_ctx, _f32p]),
    "alpine_bind_csr": (ctypes.c_int, [_c_ctx, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                       ctypes.c_void_p]),
    "alpine_fit_begin": (ctypes.c_int, [_c_ctx, ctypes.c_int, ctypes.c_void_p]),
    "alpine_batch_begin": (ctypes.c_int, [_c_ctx, ctypes.c_void_p]),
    "alpine_batch_gather": (ctypes.c_int, [_c_ctx, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "alpine_batch_scatter": (ctypes.c_int, [_c_ctx, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                            ctypes.c_int64, ctypes.c_void_p]),
    "alpine_mu_partials": (ctypes.c_int, [_c_ctx, ctypes.c_void_p]),
    "alpine_mu_apply": (ctypes.c_int, [_c_ctx, ctypes.c_int, ctypes.c_void_p]),
    "alpine_sync_w": (ctypes.c_int, [_c_ctx, ctypes.c_void_p]),
    "alpine_peer_export": (ctypes.c_int, [_c_ctx, ctypes.c_void_p]),
    "alpine_peer_import": (ctypes.c_int, [_c_ctx, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "alpine_mu_apply_peer": (ctypes.c_int, [_c_ctx, ctypes.c_int, ctypes.c_void_p]),
    "alpine_peer_disable": (ctypes.c_int, [_c_ctx]),
    "alpine_reduce_stats_offset": (ctypes.c_int64, [_c_ctx]),
    "alpine_als_block": (ctypes.c_int, [_c_ctx, ctypes.c_int, ctypes.c_void_p]),
    "alpine_als_finish": (ctypes.c_int, [_c_ctx, ctypes.c_int, ctypes.c_void_p]),
    "alpine_fit_losses": (ctypes.c_int, [_c_ctx, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                         ctypes.POINTER(ctypes.c_double), ctypes.c_void_p]),
    "alpine_eval_loss": (ctypes.c_int, [_c_ctx, ctypes.POINTER(ctypes.c_double), ctypes.c_void_p]),
    "alpine_scale": (ctypes.c_int, [_c_ctx, ctypes.c_void_p]),
    "alpine_transform": (ctypes.c_int, [_c_ctx, ctypes.c_int, ctypes.c_void_p]),
    "alpine_xh_product": (ctypes.c_int, [_c_ctx, _f32p, ctypes.c_int64, ctypes.c_void_p]),
    "alpine_wx_product": (ctypes.c_int, [_c_ctx, _f32p, ctypes.c_int64, ctypes.c_void_p]),
    "alpine_upload_rows": (ctypes.c_int, [ctypes.c_int, _f32p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]),
    "alpine_profile": (ctypes.c_int, [_c_ctx, ctypes.c_int]),
    "alpine_profile_read": (ctypes.c_int, [_c_ctx, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong)]),
    "alpine_query": (ctypes.c_int, [_c_ctx, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                    ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
}


class AlpineNativeError(RuntimeError):
    pass


def load_library(path: Optional[str] = None):
    """dlopen the C-ABI library and declare every prototype of the header."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("ALPINE_B200_LIB", LIB_PATH)
    if not os.path.isfile(p):
        raise AlpineNativeError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the MU loop)"
        )
    lib = ctypes.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _check(lib, status: int) -> None:
    if status != 0:
        msg = lib.alpine_last_error()
        raise AlpineNativeError(f"alpine_b200 native call failed (status {status}): {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load_library().alpine_launch_count())


def _stream_ptr(device: torch.device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def padded_rows(rows: int, cols: int, device, dtype=torch.float32, fill: Optional[float] = None) -> torch.Tensor:
    """(rows, cols) fp32 view whose row pitch is a multiple of 4 floats (TMA needs 16-byte pitches)."""
    ld = (cols + 3) // 4 * 4
    buf = torch.empty((rows, ld), dtype=dtype, device=device) if fill is None else torch.full(
        (rows, ld), fill, dtype=dtype, device=device)
    return buf[:, :cols]


def _host_copy_threads() -> int:
    """Staging threads of an upload: the cores of the box shared between the ranks of this node (torchrun exports
    OMP_NUM_THREADS=1, which says nothing about how many cores are idle)."""
    cores = os.cpu_count() or 1
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    return max(2, min(16, cores // max(1, local_world)))


def upload_rows(dst: torch.Tensor, src, threads: Optional[int] = None) -> None:
    """dst[:] = src for a 2-D fp32 device tensor (unit column stride) and a host array, through the library's
    uploader (csrc/host_upload.cuh): host threads stage row chunks of the pageable array into a process-wide ring of
    pinned buffers and queue the DMAs; the call returns once everything is queued and the current stream waits for
    the copies.  The GIL is released for the duration, so the caller's other threads keep running."""
    import numpy as np

    src = np.asarray(src)
    if src.dtype != np.float32:
        src = src.astype(np.float32)
    rows, cols = src.shape
    assert tuple(dst.shape) == (rows, cols) and dst.dtype == torch.float32 and dst.is_cuda
    if rows == 0 or cols == 0:
        return
    assert dst.stride(1) == 1 or cols == 1
    if src.strides[1] != 4 or src.strides[0] % 4 != 0 or src.strides[0] < cols * 4:
        src = np.ascontiguousarray(src)
    lib = load_library()
    dev = dst.device.index if dst.device.index is not None else torch.cuda.current_device()
    ld_dst = dst.stride(0) if rows > 1 else max(dst.stride(0), cols)
    ld_src = src.strides[0] // 4 if rows > 1 else cols
    _check(lib, lib.alpine_upload_rows(dev, dst.data_ptr(), ld_dst, src.ctypes.data, ld_src, rows, cols,
                                       int(threads or _host_copy_threads()), _stream_ptr(dst.device)))
    # `src` must stay alive until the staging threads are done: they are, the call returns after the last chunk was
    # copied out of it (only the DMAs out of the pinned ring are still in flight)


class Solver:
    """One device-resident shard of cells: X (cells-major), Y_i, and the factors W, H, B_i it updates in place."""

    def __init__(self, device, n_genes: int, n_cells: int, k_blocks: Sequence[int], c_cov: Sequence[int],
                 loss_type: str = "kl-divergence"):
        if not torch.cuda.is_available():
            raise AlpineNativeError("alpine_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = load_library()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise AlpineNativeError(f"device must be a CUDA device, got {device!r}")
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", self.dev_index)
        self.G, self.n = int(n_genes), int(n_cells)
        self.k_blocks = [int(k) for k in k_blocks]
        self.c_cov = [int(c) for c in c_cov]
        self.K = sum(self.k_blocks)
        self.n_cov = len(self.c_cov)
        if loss_type not in LOSS_TYPES:
            raise ValueError("loss_type must be either 'kl-divergence' or 'frobenius'.")
        kb = (ctypes.c_int * len(self.k_blocks))(*self.k_blocks)
        cc = (ctypes.c_int * max(1, self.n_cov))(*(self.c_cov or [0]))
        self._ctx = _c_ctx()
        _check(self.lib, self.lib.alpine_create(ctypes.byref(self._ctx), self.dev_index, self.G, self.n,
                                                len(self.k_blocks), kb, self.n_cov, cc, LOSS_TYPES[loss_type]))
        self._keep: dict = {}
        # per-fit workspaces come out of torch's caching allocator: no cudaMalloc / cudaFree per context
        nbytes = int(self.lib.alpine_workspace_bytes(self._ctx))
        arena = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        _check(self.lib, self.lib.alpine_bind_workspace(self._ctx, arena.data_ptr(), nbytes))
        self._keep["arena"] = arena

    # -- lifetime ----------------------------------------------------------------------------------------
    def close(self, collective: bool = True) -> None:
        """Destroy the native context.  A solver that exchanges over peer memory is closed collectively (host
        barrier first: no peer may still be reading this rank's exchange block); ``__del__`` never blocks."""
        if getattr(self, "_ctx", None) and getattr(self, "peer", False) and collective:
            import torch.distributed as dist

            torch.cuda.synchronize(self.device)
            if dist.is_initialized():
                dist.barrier(group=self._peer_group)  # no peer may still be reading this rank's exchange block
            self.peer = False
        if getattr(self, "_ctx", None):
            self.lib.alpine_destroy(self._ctx)
            self._ctx = None
            self._keep.clear()

    def __del__(self):
        try:
            self.close(collective=False)
        except Exception:
            pass

    def _stream(self) -> int:
        return _stream_ptr(self.device)

    # -- binding -----------------------------------------------------------------------------------------
    def bind_dense(self, X_cells_major: torch.Tensor) -> None:
        """X as (n_cells, n_genes) fp32 with unit column stride and a row pitch divisible by 4."""
        X = X_cells_major
        assert X.is_cuda and X.dtype == torch.float32 and X.shape == (self.n, self.G) and X.stride(1) == 1
        self._keep["X"] = X
        _check(self.lib, self.lib.alpine_bind_dense(self._ctx, X.data_ptr(), X.stride(0) if self.n > 1 else max(X.stride(0), self.G)))

    def bind_csr(self, indptr: torch.Tensor, indices: torch.Tensor, values: torch.Tensor) -> None:
        """Sparse X as device CSR over cells: indptr (n_cells + 1, int64), indices (nnz, int32 gene ids), values
        (nnz, fp32).  The library converts it into its own tile lists; the tensors are not kept."""
        assert indptr.is_cuda and indptr.dtype == torch.int64 and indptr.shape == (self.n + 1,) and indptr.is_contiguous()
        assert indices.is_cuda and indices.dtype == torch.int32 and indices.is_contiguous()
        assert values.is_cuda and values.dtype == torch.float32 and values.is_contiguous()
        assert indices.shape == values.shape and indices.dim() == 1
        self._keep.pop("X", None)
        _check(self.lib, self.lib.alpine_bind_csr(self._ctx, indptr.data_ptr(), indices.data_ptr(), values.data_ptr(),
                                                  int(values.shape[0]), self._stream()))

    def bind_labels(self, Ys: List[torch.Tensor]) -> None:
        assert len(Ys) == self.n_cov
        for i, Y in enumerate(Ys):
            assert Y.is_cuda and Y.dtype == torch.float32 and Y.shape == (self.c_cov[i], self.n) and Y.is_contiguous()
            _check(self.lib, self.lib.alpine_bind_labels(self._ctx, i, Y.data_ptr()))
        self._keep["Ys"] = list(Ys)

    def bind_factors(self, W: torch.Tensor, H: torch.Tensor, Bs: List[torch.Tensor]) -> None:
        assert W.is_cuda and W.dtype == torch.float32 and W.shape == (self.G, self.K) and W.stride(1) == 1
        assert H.is_cuda and H.dtype == torch.float32 and H.shape == (self.K, self.n) and H.stride(1) == 1
        assert len(Bs) == self.n_cov
        for i, B in enumerate(Bs):
            assert B.is_cuda and B.dtype == torch.float32 and B.is_contiguous()
            assert B.shape == (self.c_cov[i], self.k_blocks[i])
        arr = (ctypes.c_void_p * max(1, self.n_cov))(*([B.data_ptr() for B in Bs] or [None]))
        self._keep.update(W=W, H=H, Bs=list(Bs))
        ldH = H.stride(0) if self.K > 1 else max(H.stride(0), self.n)
        ldW = W.stride(0) if self.G > 1 else max(W.stride(0), self.K)
        _check(self.lib, self.lib.alpine_bind_factors(self._ctx, W.data_ptr(), ldW, H.data_ptr(), ldH, arr))

    def set_hparams(self, lam: Sequence[float], alpha_W: float, l1_ratio_W: float, orth_W: float, eps: float) -> None:
        arr = (ctypes.c_double * max(1, self.n_cov))(*([float(v) for v in lam] or [0.0]))
        _check(self.lib, self.lib.alpine_set_hparams(self._ctx, arr, float(alpha_W), float(l1_ratio_W), float(orth_W),
                                                     float(eps)))

    # -- peer exchange over NVLink (one process per GPU) ----------------------------------------------------
    def enable_peer_exchange(self, group=None) -> bool:
        """Collective over the ranks of ``group``: share every rank's exchange block through CUDA IPC so that the W
        update exchanges over peer memory (``mu_apply_peer``) instead of an all-reduce.  Must be called before
        ``reduce_buffer`` / ``fit_begin``.  Returns False (on every rank) when some rank could not set it up; the
        caller then stays on the all-reduce path."""
        import torch.distributed as dist

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        handle = (ctypes.c_ubyte * 64)()
        ok = self.lib.alpine_peer_export(self._ctx, ctypes.cast(handle, ctypes.c_void_p)) == 0
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=self.device)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine, group=group)
        flag = torch.tensor([1 if ok else 0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 1:
            blob = b"".join(bytes(g.cpu().numpy().tobytes()) for g in gathered)
            buf = ctypes.create_string_buffer(blob, len(blob))
            ok = self.lib.alpine_peer_import(self._ctx, rank, world, ctypes.cast(buf, ctypes.c_void_p)) == 0
            flag = torch.tensor([1 if ok else 0], device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.peer = int(flag.item()) == 1
        if not self.peer:
            self.lib.alpine_peer_disable(self._ctx)  # every rank falls back to the all-reduce path together
        self._peer_group = group
        dist.barrier(group=group)
        return self.peer

    def mu_apply_peer(self, it: int) -> None:
        _check(self.lib, self.lib.alpine_mu_apply_peer(self._ctx, int(it), self._stream()))

    def reduce_buffer(self) -> torch.Tensor:
        """Allocate and bind the per-iteration all-reduce payload [X H^T | H H^T | rowsum(H) | B statistics]."""
        if getattr(self, "peer", False):
            raise AlpineNativeError("this solver exchanges over peer memory; there is no caller-owned reduce buffer")
        if "reduce" not in self._keep:
            size = int(self.lib.alpine_reduce_buffer_size(self._ctx))
            buf = torch.zeros(size, dtype=torch.float32, device=self.device)
            _check(self.lib, self.lib.alpine_bind_reduce_buffer(self._ctx, buf.data_ptr()))
            self._keep["reduce"] = buf
        return self._keep["reduce"]

    # -- the loop ----------------------------------------------------------------------------------------
    def fit_begin(self, max_iter: int) -> None:
        _check(self.lib, self.lib.alpine_fit_begin(self._ctx, int(max_iter), self._stream()))

    def batch_begin(self) -> None:
        _check(self.lib, self.lib.alpine_batch_begin(self._ctx, self._stream()))

    def batch_gather(self, X_all: Optional[torch.Tensor], H_all: torch.Tensor, Ys_all: List[torch.Tensor],
                     idx: torch.Tensor) -> None:
        """Cells ``idx`` of the full-data arrays into the bound batch arrays, the rest of the context zero-filled
        (one launch; main.py:593-595).  ``X_all``: cells-major (n_all, n_genes), None for a CSR context."""
        n_all = H_all.shape[1]
        assert idx.is_cuda and idx.dtype == torch.int64 and idx.dim() == 1 and idx.is_contiguous() and len(idx) <= self.n
        assert H_all.is_cuda and H_all.dtype == torch.float32 and H_all.shape[0] == self.K and H_all.stride(1) == 1
        assert len(Ys_all) == self.n_cov
        for i, Y in enumerate(Ys_all):
            assert Y.is_cuda and Y.dtype == torch.float32 and Y.shape == (self.c_cov[i], n_all) and Y.is_contiguous()
        Xb = self._keep.get("X")
        if Xb is not None:
            assert X_all is not None and X_all.is_cuda and X_all.dtype == torch.float32 and X_all.stride(1) == 1
            assert X_all.shape == (n_all, self.G)
        ya = (ctypes.c_void_p * max(1, self.n_cov))(*([Y.data_ptr() for Y in Ys_all] or [None]))
        yb = (ctypes.c_void_p * max(1, self.n_cov))(*([Y.data_ptr() for Y in self._keep.get("Ys", [])] or [None]))
        ldH = H_all.stride(0) if self.K > 1 else max(H_all.stride(0), n_all)
        _check(self.lib, self.lib.alpine_batch_gather(
            self._ctx, X_all.data_ptr() if Xb is not None else None,
            (X_all.stride(0) if n_all > 1 else max(X_all.stride(0), self.G)) if Xb is not None else 0,
            Xb.data_ptr() if Xb is not None else None, H_all.data_ptr(), ldH, ya, yb, n_all, idx.data_ptr(), len(idx),
            self._stream()))

    def batch_scatter(self, H_all: torch.Tensor, idx: torch.Tensor) -> None:
        """``H_all[:, idx] = H_batch[:, :len(idx)]`` (main.py:662), one launch."""
        n_all = H_all.shape[1]
        assert idx.is_cuda and idx.dtype == torch.int64 and idx.dim() == 1 and idx.is_contiguous() and len(idx) <= self.n
        assert H_all.is_cuda and H_all.dtype == torch.float32 and H_all.shape[0] == self.K and H_all.stride(1) == 1
        ldH = H_all.stride(0) if self.K > 1 else max(H_all.stride(0), n_all)
        _check(self.lib, self.lib.alpine_batch_scatter(self._ctx, H_all.data_ptr(), ldH, n_all, idx.data_ptr(), len(idx),
                                                       self._stream()))

    def mu_partials(self) -> None:
        _check(self.lib, self.lib.alpine_mu_partials(self._ctx, self._stream()))

    def mu_apply(self, it: int) -> None:
        _check(self.lib, self.lib.alpine_mu_apply(self._ctx, int(it), self._stream()))

    def sync_w(self) -> None:
        """Refresh the bound row-major W from the library's W^T master copy (the updates only touch the latter)."""
        _check(self.lib, self.lib.alpine_sync_w(self._ctx, self._stream()))

    # -- block Gauss-Seidel sweep (use_als=True, main.py:523-588) ------------------------------------------
    @property
    def n_blocks(self) -> int:
        return len(self.k_blocks)

    def gram_view(self) -> torch.Tensor:
        """The H H^T part (K*K floats) of the reduce buffer: the per-block exchange of the ALS sweep."""
        buf = self.reduce_buffer()
        o = int(self.lib.alpine_reduce_stats_offset(self._ctx))
        return buf[o:o + self.K * self.K]

    def als_block(self, b: int) -> None:
        _check(self.lib, self.lib.alpine_als_block(self._ctx, int(b), self._stream()))

    def als_finish(self, it: int) -> None:
        _check(self.lib, self.lib.alpine_als_finish(self._ctx, int(it), self._stream()))

    def losses(self, n_iter: int):
        """(||X||_F^2, rows[n_iter][2 + n_cov]) of this shard; synchronises the stream."""
        import numpy as np

        xn = ctypes.c_double(0.0)
        rows = np.zeros((max(n_iter, 1), 2 + self.n_cov), dtype=np.float64)
        _check(self.lib, self.lib.alpine_fit_losses(self._ctx, int(n_iter), ctypes.byref(xn),
                                                    rows.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                                    self._stream()))
        return float(xn.value), rows[:n_iter]

    def eval_loss(self):
        """[tr(W^T X H^T), tr(W^T W H H^T), pred_0, ...] of the bound factors on this shard (fp64); synchronises."""
        import numpy as np

        terms = np.zeros(2 + self.n_cov, dtype=np.float64)
        _check(self.lib, self.lib.alpine_eval_loss(self._ctx, terms.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                                   self._stream()))
        return terms

    def scale(self) -> None:
        _check(self.lib, self.lib.alpine_scale(self._ctx, self._stream()))

    def transform(self, n_iter: int) -> None:
        _check(self.lib, self.lib.alpine_transform(self._ctx, int(n_iter), self._stream()))

    # -- the two contractions on their own -----------------------------------------------------------------
    def xh_product(self) -> torch.Tensor:
        """(K, G): sum_j X[g][j] H[k][j] with the bound X and H."""
        out = padded_rows(self.K, self.G, self.device)
        _check(self.lib, self.lib.alpine_xh_product(self._ctx, out.data_ptr(), out.stride(0), self._stream()))
        return out

    def wx_product(self) -> torch.Tensor:
        """(K, n): sum_g W[g][k] X[g][j] with the bound X and W."""
        out = padded_rows(self.K, self.n, self.device)
        _check(self.lib, self.lib.alpine_wx_product(self._ctx, out.data_ptr(), out.stride(0), self._stream()))
        return out

    def profile(self, enable: bool) -> None:
        _check(self.lib, self.lib.alpine_profile(self._ctx, 1 if enable else 0))

    def profile_read(self):
        """(total ms, launches) of the contraction kernel since profile(True); CUDA events on the launch stream."""
        ms, cnt = ctypes.c_double(0.0), ctypes.c_longlong(0)
        _check(self.lib, self.lib.alpine_profile_read(self._ctx, ctypes.byref(ms), ctypes.byref(cnt)))
        return float(ms.value), int(cnt.value)

    def query(self) -> dict:
        a, b, c, d = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _check(self.lib, self.lib.alpine_query(self._ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c),
                                               ctypes.byref(d)))
        return {"num_sms": a.value, "gemm_grid": b.value, "smem_stages": c.value, "k_padded": d.value}
