"""CPU tests of the host-side sharding logic (MUEngine) with torch.distributed/gloo, world_size 2.

The CUDA solver is replaced by tests/np_shard_solver.py (same per-shard dataflow in NumPy); the result of the
2-rank run must match the single-process oracle on the full matrix, which checks the column partition, the single
per-iteration all-reduce of the packed buffer and the loss aggregation.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alpine_b200.engine import MUEngine, shard_bounds
from oracle import alpine_oracle as orc
from tests.helpers import CASE_KW, hp_of, inputs_of, load_golden, rel_fro
from tests.np_shard_solver import NumpyShardSolver


def test_shard_bounds_partition():
    for n in (1, 7, 100, 100001):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for (a, b), (c, d) in zip(blocks[:-1], blocks[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run_rank(rank, world, port, name, n_iter, out_dir, peer=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load_golden(name)
        hp = hp_of(name)
        X, Ys, st = inputs_of(g)
        lo, hi = shard_bounds(X.shape[1], world, rank)
        solver = NumpyShardSolver(np.ascontiguousarray(X[:, lo:hi]), [np.ascontiguousarray(y[:, lo:hi]) for y in Ys],
                                  st.W.copy(), np.ascontiguousarray(st.H[:, lo:hi]), [b.copy() for b in st.Bs],
                                  st.blocks, hp)
        if peer:
            assert solver.enable_peer_exchange()
        engine = MUEngine(solver, hp.lam, use_als=CASE_KW[name].get("use_als", False))
        assert (engine.rank, engine.world) == (rank, world)
        hist = engine.run(n_iter)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), W=solver.W, H=solver.H, hist=hist, lo=lo, hi=hi,
                 **{f"B{i}": b for i, b in enumerate(solver.Bs)})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["kl_reg_nan", "frob_reg", "als_reg"])
def test_two_rank_run_matches_oracle(name, tmp_path):
    world, n_iter = 2, 5
    mp.spawn(_run_rank, args=(world, _free_port(), name, n_iter, str(tmp_path)), nprocs=world, join=True)
    g = load_golden(name)
    hp = hp_of(name)
    X, Ys, st = inputs_of(g)
    hist_ref, _ = orc.fit_loop(X, Ys, st, hp, n_iter, use_als=CASE_KW[name].get("use_als", False))
    ref64 = orc.compute_loss(X, Ys, st, hp, dtype=np.float64)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    H = np.concatenate([p["H"] for p in parts], axis=1)
    assert rel_fro(H, st.H) < 2e-5
    for p in parts:  # W and B are replicas: identical on every rank
        assert rel_fro(p["W"], st.W) < 2e-5
        np.testing.assert_array_equal(p["W"], parts[0]["W"])
        for i in range(len(Ys)):
            assert rel_fro(p[f"B{i}"], st.Bs[i]) < 2e-5
        np.testing.assert_array_equal(p["hist"], parts[0]["hist"])  # every rank holds the global loss history
    hist = parts[0]["hist"]
    assert hist.shape == (n_iter, 2 + len(Ys))
    assert abs(hist[-1, 1] - ref64[1]) / ref64[1] < 1e-4          # trace-identity reconstruction loss
    np.testing.assert_allclose(hist[-1, 2:], ref64[2:], rtol=1e-3, atol=1e-6 * X.shape[1])
    np.testing.assert_allclose(hist[-1, 0], hist[-1, 1] + sum(l * p for l, p in zip(hp.lam, hist[-1, 2:])), rtol=1e-12)


@pytest.mark.parametrize("world", [2, 3])
def test_peer_exchange_mode_matches_oracle(world, tmp_path):
    """MUEngine's peer branch (mu_partials + mu_apply_peer, no all-reduce of the buffer) with the stand-in solver's
    gene-slice exchange; 3 ranks make the 64-gene-tile slices ragged."""
    name, n_iter = "kl_reg_nan", 4
    mp.spawn(_run_rank, args=(world, _free_port(), name, n_iter, str(tmp_path), True), nprocs=world, join=True)
    g = load_golden(name)
    hp = hp_of(name)
    X, Ys, st = inputs_of(g)
    orc.fit_loop(X, Ys, st, hp, n_iter)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert rel_fro(np.concatenate([p["H"] for p in parts], axis=1), st.H) < 2e-5
    for p in parts:
        assert rel_fro(p["W"], st.W) < 2e-5
        np.testing.assert_array_equal(p["W"], parts[0]["W"])     # bit-identical replicas
        np.testing.assert_array_equal(p["hist"], parts[0]["hist"])
    ref64 = orc.compute_loss(X, Ys, st, hp, dtype=np.float64)
    assert abs(parts[0]["hist"][-1, 1] - ref64[1]) / ref64[1] < 1e-4


def test_single_process_engine_matches_oracle():
    name = "kl_basic"
    g = load_golden(name)
    hp = hp_of(name)
    X, Ys, st = inputs_of(g)
    solver = NumpyShardSolver(X.copy(), [y.copy() for y in Ys], st.W.copy(), st.H.copy(), [b.copy() for b in st.Bs],
                              st.blocks, hp)
    hist = MUEngine(solver, hp.lam).run(10)
    assert rel_fro(solver.W, g["W_it10"]) < 2e-5
    assert rel_fro(solver.H, g["H_it10"]) < 2e-5
    assert abs(hist[-1, 1] - float(g["final_recon_fp64"])) / float(g["final_recon_fp64"]) < 1e-4
