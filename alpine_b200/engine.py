"""Host-side driver of the MU loop: one process per GPU, cells sharded across ranks.

The reference runs the whole loop on one device (``ALPINE._fit``, main.py:486-676).  Here every rank owns a
column block of X / H / Y (cells) and a replica of W / B; per iteration the only exchange is ONE all-reduce (sum)
of the packed buffer ``[X H^T | H H^T | rowsum(H) | B statistics]``, issued through
``torch.distributed`` (NCCL on the GPUs; gloo in the CPU tests of this logic).  After it every rank applies the
identical W and B updates and updates its own H block; the loss terms are additive across ranks and are summed
once, after the last iteration.

``MUEngine`` is independent of the CUDA library: it drives any object with the ``ShardSolver`` methods, which is
how the sharding logic is tested on CPU with world_size 2.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Protocol, Sequence, Tuple

import numpy as np
import torch


def shard_bounds(n_cells: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced column block [lo, hi) of rank ``rank``; blocks differ in size by at most one cell."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, extra = divmod(int(n_cells), int(world_size))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


class ShardSolver(Protocol):
    n_cov: int

    def reduce_buffer(self) -> torch.Tensor: ...
    def fit_begin(self, max_iter: int) -> None: ...
    def mu_partials(self) -> None: ...
    def mu_apply(self, it: int) -> None: ...
    # block Gauss-Seidel sweep only: n_blocks, als_block(b), gram_view(), als_finish(it)
    def losses(self, n_iter: int) -> Tuple[float, np.ndarray]: ...


def dist_info(group=None) -> Tuple[int, int]:
    """(rank, world_size) of the active process group, (0, 1) when torch.distributed is not initialised."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class MUEngine:
    """Runs ``max_iter`` full-batch MU iterations on this rank's shard and returns the global loss history."""

    def __init__(self, solver: ShardSolver, lam: Sequence[float], group=None, use_als: bool = False):
        self.solver = solver
        self.lam = [float(v) for v in lam]
        self.group = group
        self.use_als = bool(use_als)
        self.rank, self.world = dist_info(group)

    def _all_reduce(self, t: torch.Tensor) -> None:
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def run(self, max_iter: int, on_iter: Optional[Callable[[int], None]] = None) -> np.ndarray:
        """Returns rows ``[total, reconstruction, prediction_0, ...]`` per iteration (main.py:750-753)."""
        self.begin(max_iter)
        for it in range(max_iter):
            self.step(it)
            if on_iter is not None:
                on_iter(it)
        return self.collect_losses(max_iter)

    def begin(self, max_iter: int) -> None:
        """||X||^2 and this shard's statistics of the initial H / B (left in the reduce buffer)."""
        self.peer = bool(getattr(self.solver, "peer", False)) and self.world > 1 and not self.use_als
        # a caller-owned exchange buffer only where an all-reduce follows: on one GPU the W update takes its numerator
        # straight from the contraction's partial sums (one kernel and one G x K round trip less per iteration)
        need_buf = self.world > 1 and not self.peer
        self._buf = self.solver.reduce_buffer() if need_buf else None
        self.solver.fit_begin(max_iter)

    def step(self, it: int) -> None:
        """One full-batch MU iteration (asynchronous on the current stream)."""
        s = self.solver
        if self.use_als:
            return self._als_step(it)
        if self.peer:
            # exchange over NVLink peer memory inside the W-update kernels (csrc/peer_exchange.cuh): no collective
            s.mu_partials()
            s.mu_apply_peer(it)
            return
        s.mu_partials()        # X H^T of this shard into the reduce buffer (main.py:596)
        # The iteration's only data-path collective.  The statistics part [H H^T | rowsum(H) | B statistics]
        # still holds this shard's values (written by fit_begin / the previous mu_apply), so one message
        # carries everything the W and B updates need.
        if self._buf is not None:
            self._all_reduce(self._buf)
        s.mu_apply(it)         # W, B, H updates + loss terms (main.py:597-663, 726-753)

    def _als_step(self, it: int) -> None:
        """One block Gauss-Seidel sweep (use_als=True, main.py:523-588): blocks in order, each W_b, B_b, H_b.

        X H_b^T is needed with the H_b the block starts from, which is the H of the iteration start for every b, so
        one sweep of X (and one all-reduce of the packed buffer) serves all blocks; after a block's H update only
        H H^T changes for the blocks that follow: K*K floats are exchanged per block."""
        s = self.solver
        s.mu_partials()
        if self._buf is not None:
            self._all_reduce(self._buf)
        n_blocks = s.n_blocks
        for b in range(n_blocks):
            s.als_block(b)
            if b + 1 < n_blocks and self.world > 1:
                self._all_reduce(s.gram_view())
        s.als_finish(it)

    def collect_losses(self, n_iter: int) -> np.ndarray:
        xn, rows = self.solver.losses(n_iter)
        n_cov = rows.shape[1] - 2
        packed = np.concatenate([[xn], rows.reshape(-1)]).astype(np.float64)
        if self.world > 1:
            dev = getattr(self.solver, "device", None) or self.solver.reduce_buffer().device
            t = torch.from_numpy(packed).to(dev)
            self._all_reduce(t)
            packed = t.cpu().numpy()
        xn = float(packed[0])
        rows = packed[1:].reshape(n_iter, 2 + n_cov)
        out = np.zeros((n_iter, 2 + n_cov), dtype=np.float64)
        out[:, 1] = xn - 2.0 * rows[:, 0] + rows[:, 1]  # ||X||^2 - 2 tr(W^T X H^T) + tr(W^T W H H^T)
        out[:, 2:] = rows[:, 2:]
        out[:, 0] = out[:, 1] + sum(self.lam[i] * out[:, 2 + i] for i in range(n_cov))
        return out
