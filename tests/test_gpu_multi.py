"""Multi-GPU parity under the driver's `-m gpu` run: skipped on a one-GPU box, otherwise launches one process per
visible GPU (torch.distributed.run, NCCL) and requires what tools/dist_check.py requires: the cell-sharded fit --
through the NCCL all-reduce AND through the NVLink peer-memory exchange kernels -- equals the single-GPU fit to 1e-5
(W, B, every rank's H block, reconstruction loss), with W bit-identical across ranks, at K = 100 with ragged shards.
"""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _launch(n_gpus, extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_gpus}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + (os.getpid() % 400)),
           os.path.join(ROOT, "tools", "dist_check.py")] + extra
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("dist_check ")]
    assert out.returncode == 0 and lines, (out.returncode, out.stdout[-2000:], out.stderr[-2000:])
    return json.loads(lines[-1][len("dist_check "):])


@pytest.mark.parametrize("shape", [
    dict(cells=6001, genes=1500, iters=8),     # 750-cell shards at 8 GPUs: ragged, ~3 super-tiles per rank
    dict(cells=20003, genes=2600, iters=4),    # several stream-K pieces / slots per rank
])
def test_cell_sharded_fit_equals_single_gpu_fit(shape):
    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip("needs >= 2 GPUs (one process per GPU)")
    res = _launch(n_gpus, ["--cells", str(shape["cells"]), "--genes", str(shape["genes"]), "--iters", str(shape["iters"])])
    assert res["ok"], res
    assert res["world"] == n_gpus
    for mode in ("nccl", "peer"):
        assert res[mode]["W_bit_identical_across_ranks"], res
        assert max(res[mode][k] for k in ("W", "H", "B", "recon_loss")) < 1e-5, res
    assert res["peer"]["peer_exchange_active"], res


def test_minibatch_epochs_under_cell_sharding_equal_single_gpu():
    """batch_size < n_cells with the cells sharded (main.py:509-521 on every rank's part of the batch + the usual
    all-reduce): random and weighted samplers and the block-wise sweep, against the same fit on one GPU."""
    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip("needs >= 2 GPUs (one process per GPU)")
    res = _launch(n_gpus, ["--minibatch"])
    assert res["ok"] and res["world"] == n_gpus, res
