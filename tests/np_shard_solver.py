"""NumPy stand-in for ``alpine_b200._native.Solver`` (TEST INFRASTRUCTURE, CPU only).

It implements the ``ShardSolver`` protocol of ``alpine_b200.engine`` with the same per-shard dataflow as the CUDA
library (X H^T partials -> packed reduce buffer -> W / B / H updates -> statistics of the new H), so that the
host-side sharding logic (``MUEngine``: one all-reduce per iteration, loss aggregation) can be exercised with
``torch.distributed``'s gloo backend and world_size 2 on a machine without a GPU.
"""
from __future__ import annotations

import numpy as np
import torch


class NumpyShardSolver:
    def __init__(self, X_gc, Ys, W, H, Bs, blocks, hp):
        """X_gc: genes x local cells; Ys: c_i x local cells; W: G x K (replica); H: K x local cells."""
        self.X, self.Ys, self.W, self.H, self.Bs = X_gc, Ys, W, H, Bs
        self.blocks, self.hp = blocks, hp
        self.K = W.shape[1]
        self.G = W.shape[0]
        self.n_cov = len(Ys)
        self.sl, s = [], 0
        for k in blocks:
            self.sl.append(slice(s, s + k))
            s += k
        self.q_sizes = [Ys[i].shape[0] * blocks[i] for i in range(self.n_cov)]
        size = self.K * self.G + self.K * self.K + self.K + sum(self.q_sizes)
        self._buf = torch.zeros(size, dtype=torch.float32)
        self.rows = []

    # layout [Pt (K x G) | S (K x K) | hsum (K) | Q_i ...]
    def _views(self):
        b = self._buf.numpy()
        o = 0
        Pt = b[o:o + self.K * self.G].reshape(self.K, self.G); o += self.K * self.G
        S = b[o:o + self.K * self.K].reshape(self.K, self.K); o += self.K * self.K
        hs = b[o:o + self.K]; o += self.K
        Qs = []
        for i in range(self.n_cov):
            Qs.append(b[o:o + self.q_sizes[i]].reshape(self.Ys[i].shape[0], self.blocks[i])); o += self.q_sizes[i]
        return Pt, S, hs, Qs

    def reduce_buffer(self):
        return self._buf

    def _stats(self):
        Pt, S, hs, Qs = self._views()
        H = self.H
        S[...] = H @ H.T
        hs[...] = H.sum(axis=1)
        eps = np.float32(self.hp.eps)
        pred = []
        for i in range(self.n_cov):
            Hi, B, Y = H[self.sl[i]], self.Bs[i], self.Ys[i]
            yh = B @ Hi
            if self.hp.loss_type == "kl-divergence":
                yc = np.maximum(yh, eps)
                Qs[i][...] = (Y / yc) @ Hi.T
                pred.append(float(np.sum(Y * np.log(np.maximum(Y / yc, eps)) - Y + yc, dtype=np.float64)))
            else:
                Qs[i][...] = Y @ Hi.T
                pred.append(float(np.sum((Y - yh).astype(np.float64) ** 2)))
        return pred

    def fit_begin(self, max_iter):
        self.xn = float(np.sum(self.X.astype(np.float64) ** 2))
        self.rows = []
        self._stats()

    def mu_partials(self):
        Pt, _, _, _ = self._views()
        Pt[...] = self.H @ self.X.T

    def mu_apply(self, it):
        hp = self.hp
        Pt, S, hs, Qs = self._views()
        eps = np.float32(hp.eps)
        W = self.W
        c1, c2 = np.float32((1 - hp.l1_ratio_W) * hp.alpha_W), np.float32(hp.l1_ratio_W * hp.alpha_W)
        den = 2 * (W @ S) + c1 * W + np.float32(hp.orth_W) * (W.sum(axis=1, keepdims=True) - W) + c2
        W *= (2 * Pt.T) / np.maximum(den, eps)
        self._apply_after_w(it)

    def _apply_after_w(self, it):
        """B updates, H update, loss terms and statistics, given the new W (shared by both exchange modes)."""
        hp = self.hp
        Pt, S, hs, Qs = self._views()
        eps = np.float32(hp.eps)
        W = self.W
        for i in range(self.n_cov):
            B, lam = self.Bs[i], np.float32(hp.lam[i])
            if hp.loss_type == "kl-divergence":
                B *= (lam * Qs[i]) / np.maximum(lam * hs[self.sl[i]][None, :], eps)
            else:
                B *= (2 * Qs[i]) / np.maximum((2 * B) @ S[self.sl[i], self.sl[i]], eps)
        T = W.T @ W
        A = W.T @ self.X
        H = self.H
        num = 2 * A
        den = 2 * (T @ H)
        for i in range(self.n_cov):
            B, lam, Hi, Y = self.Bs[i], np.float32(hp.lam[i]), H[self.sl[i]], self.Ys[i]
            if hp.loss_type == "kl-divergence":
                num[self.sl[i]] += (lam * B.T) @ (Y / np.maximum(B @ Hi, eps))
                den[self.sl[i]] += ((lam * B.T) @ np.ones_like(Y))
            else:
                num[self.sl[i]] += (2 * lam * B.T) @ Y
                den[self.sl[i]] += (2 * lam * B.T) @ (B @ Hi)
        H *= num / np.maximum(den, eps)
        t1 = float(np.sum(A.astype(np.float64) * H.astype(np.float64)))
        pred = self._stats()
        _, S2, _, _ = self._views()
        t2 = float(np.sum(T.astype(np.float64) * S2.astype(np.float64)))
        self.rows.append([t1, t2] + pred)

    # ---- peer-exchange mode (same dataflow as csrc/peer_exchange.cuh, with gloo standing in for NVLink loads / stores):
    # every rank sums the ranks' partials only for ITS gene slice (whole 64-gene tiles), updates that slice of W, and
    # the slices are gathered; the small statistics are summed in rank order on every rank.
    peer = False

    def enable_peer_exchange(self, group=None):
        import torch.distributed as dist

        self.peer, self._group = True, group
        self._rank, self._world = dist.get_rank(group), dist.get_world_size(group)
        return True

    def mu_apply_peer(self, it):
        import torch.distributed as dist

        bufs = [torch.empty_like(self._buf) for _ in range(self._world)]
        dist.all_gather(bufs, self._buf, group=self._group)          # "peer loads" of every rank's exchange block
        small0 = self.K * self.G
        tiles = (self.G + 63) // 64
        g0 = tiles * self._rank // self._world * 64
        g1 = min(self.G, tiles * (self._rank + 1) // self._world * 64)
        mine = self._buf.numpy()
        small = np.zeros(mine.size - small0, dtype=np.float32)
        for b in bufs:                                               # rank order: identical sums on every rank
            small += b.numpy()[small0:]
        P_slice = np.zeros((self.K, g1 - g0), dtype=np.float32)
        for b in bufs:
            P_slice += b.numpy()[:small0].reshape(self.K, self.G)[:, g0:g1]
        local_small = mine[small0:].copy()
        mine[small0:] = small                                        # the updates below read the summed statistics
        Pt, S, hs, Qs = self._views()
        hp = self.hp
        eps = np.float32(hp.eps)
        W = self.W
        c1, c2 = np.float32((1 - hp.l1_ratio_W) * hp.alpha_W), np.float32(hp.l1_ratio_W * hp.alpha_W)
        Ws = W[g0:g1]
        den = 2 * (Ws @ S) + c1 * Ws + np.float32(hp.orth_W) * (Ws.sum(axis=1, keepdims=True) - Ws) + c2
        new_slice = torch.from_numpy(np.ascontiguousarray(Ws * ((2 * P_slice.T) / np.maximum(den, eps))))
        sizes = [min(self.G, tiles * (r + 1) // self._world * 64) - tiles * r // self._world * 64 for r in range(self._world)]
        parts = [torch.empty((sz, self.K), dtype=torch.float32) for sz in sizes]
        dist.all_gather(parts, new_slice, group=self._group) if len(set(sizes)) == 1 else self._gather_ragged(parts, new_slice)
        W[...] = np.concatenate([p.numpy() for p in parts], axis=0)  # "peer stores" of every rank's new slice
        Pt[...] = 0  # the numerator was consumed from the gathered blocks; mu_apply's W part must not run again
        self._apply_after_w(it)
        # like the CUDA path, the exchange block keeps THIS rank's partial statistics for the next iteration: they
        # were refreshed by _apply_after_w (via _stats), nothing to restore
        del local_small

    def _gather_ragged(self, parts, mine):
        import torch.distributed as dist

        for r, p in enumerate(parts):
            if r == self._rank:
                p.copy_(mine)
            dist.broadcast(p, src=r, group=self._group)

    # ---- block Gauss-Seidel sweep (same split as alpine_als_block / alpine_als_finish)
    @property
    def n_blocks(self):
        return len(self.blocks)

    def gram_view(self):
        o = self.K * self.G
        return self._buf[o:o + self.K * self.K]

    def als_block(self, b):
        hp = self.hp
        Pt, S, hs, Qs = self._views()
        eps = np.float32(hp.eps)
        sl = self.sl[b]
        W, H = self.W, self.H
        c1, c2 = np.float32((1 - hp.l1_ratio_W) * hp.alpha_W), np.float32(hp.l1_ratio_W * hp.alpha_W)
        Wb = W[:, sl]
        den = 2 * (W @ S[:, sl]) + c1 * Wb + np.float32(hp.orth_W) * (Wb.sum(axis=1, keepdims=True) - Wb) + c2
        W[:, sl] = Wb * ((2 * Pt[sl].T) / np.maximum(den, eps))
        if b < self.n_cov:
            B, lam = self.Bs[b], np.float32(hp.lam[b])
            if hp.loss_type == "kl-divergence":
                B *= (lam * Qs[b]) / np.maximum(lam * hs[sl][None, :], eps)
            else:
                B *= (2 * Qs[b]) / np.maximum((2 * B) @ S[sl, sl], eps)
        T = W.T @ W
        if not hasattr(self, "A"):
            self.A = np.zeros_like(H)
        self.A[sl] = W[:, sl].T @ self.X
        num = 2 * self.A[sl]
        den = 2 * (T[sl] @ H)
        if b < self.n_cov:
            B, lam, Hi, Y = self.Bs[b], np.float32(hp.lam[b]), H[sl], self.Ys[b]
            if hp.loss_type == "kl-divergence":
                num = num + (lam * B.T) @ (Y / np.maximum(B @ Hi, eps))
                den = den + (lam * B.T) @ np.ones_like(Y)
            else:
                num = num + (2 * lam * B.T) @ Y
                den = den + (2 * lam * B.T) @ (B @ Hi)
        H[sl] = H[sl] * (num / np.maximum(den, eps))
        if b + 1 < len(self.blocks):
            S[...] = H @ H.T

    def als_finish(self, it):
        t1 = float(np.sum(self.A.astype(np.float64) * self.H.astype(np.float64)))
        pred = self._stats()
        _, S2, _, _ = self._views()
        T = self.W.T @ self.W
        t2 = float(np.sum(T.astype(np.float64) * S2.astype(np.float64)))
        self.rows.append([t1, t2] + pred)

    def losses(self, n_iter):
        return self.xn, np.asarray(self.rows[:n_iter], dtype=np.float64).reshape(n_iter, 2 + self.n_cov)
