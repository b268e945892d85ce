"""Run under torchrun with N >= 2 GPUs: cell-sharded fit == single-GPU fit (same inputs), within fp tolerance.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpine_b200 import _native  # noqa: E402
from alpine_b200.engine import MUEngine, shard_bounds  # noqa: E402


def build(dev, X, Ys, W, H, Bs, blocks, cats, kw):
    s = _native.Solver(dev, X.shape[1], X.shape[0], blocks, cats)
    s.bind_dense(X)
    s.bind_labels(Ys)
    s.bind_factors(W, H, Bs)
    s.set_hparams(kw["lam"], kw["alpha_W"], kw["l1_ratio_W"], kw["orth_W"], 1e-6)
    return s


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, G, blocks, cats = 6001, 1500, [4, 3, 20], [3, 4]
    kw = dict(lam=[1e3, 5e2], alpha_W=0.5, l1_ratio_W=0.5, orth_W=0.2)
    K = sum(blocks)
    g = torch.Generator(device="cpu").manual_seed(0)
    Xh = torch.rand((n, G), generator=g).pow(3.0)
    Wh = torch.rand((G, K), generator=g).clamp(min=1e-6)
    Hh = torch.rand((K, n), generator=g).clamp(min=1e-6)
    Bh = [torch.rand((c, k), generator=g).clamp(min=1e-6) for c, k in zip(cats, blocks)]
    Yh = [torch.nn.functional.one_hot(torch.randint(0, c, (n,), generator=g), c).T.float().contiguous() for c in cats]
    n_iter = 8

    def run(lo, hi, use_dist, peer=False):
        X = _native.padded_rows(hi - lo, G, dev)
        X.copy_(Xh[lo:hi])
        H = _native.padded_rows(K, hi - lo, dev)
        H.copy_(Hh[:, lo:hi])
        W = Wh.clone().to(dev)
        Bs = [b.clone().to(dev) for b in Bh]
        Ys = [y[:, lo:hi].contiguous().to(dev) for y in Yh]
        s = build(dev, X, Ys, W, H, Bs, blocks, cats, kw)
        if peer:
            assert s.enable_peer_exchange(), "CUDA IPC peer exchange could not be set up"
        eng = MUEngine(s, kw["lam"])
        if not use_dist:
            eng.world = 1
        hist = eng.run(n_iter)
        out = (W.cpu().numpy(), H.cpu().numpy(), [b.cpu().numpy() for b in Bs], hist)
        s.close()
        return out

    lo, hi = shard_bounds(n, world, rank)
    Wd, Hd, Bd, hist_d = run(lo, hi, True)
    # the same sharded fit exchanging over NVLink peer memory instead of the NCCL all-reduce
    Wp, Hp, Bp, hist_p = run(lo, hi, True, peer=True)
    gathered = [None] * world
    dist.all_gather_object(gathered, Wp.tobytes())
    same_w = all(g == gathered[0] for g in gathered)  # W must be bit-identical on every rank
    ok = True
    if rank == 0:
        W1, H1, B1, hist_1 = run(0, n, False)

        def rel(a, b):
            return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))

        eW, eH = rel(Wd, W1), rel(Hd, H1[:, lo:hi])
        eB = max(rel(a, b) for a, b in zip(Bd, B1))
        col_err = np.max(np.abs(hist_d - hist_1) / np.abs(hist_1), axis=0)
        # prediction (KL) terms cancel element-wise (y log(y/yh) - y + yh ~ (yh-1)^2/2): compare them absolutely
        pred_abs = float(np.max(np.abs(hist_d[:, 2:] - hist_1[:, 2:]))) / n
        print(f"dist_check world={world}: W {eW:.2e}  H(block0) {eH:.2e}  B {eB:.2e}  loss rel err per column "
              f"{np.array2string(col_err, precision=2)}  pred abs err / n {pred_abs:.2e}")
        print("  last rows:", hist_d[-1], hist_1[-1])
        ok = eW < 1e-5 and eH < 1e-5 and eB < 1e-5 and col_err[1] < 1e-5 and pred_abs < 1e-6
        ePW, ePH = rel(Wp, W1), rel(Hp, H1[:, lo:hi])
        ePB = max(rel(a, b) for a, b in zip(Bp, B1))
        p_err = np.max(np.abs(hist_p[:, :2] - hist_1[:, :2]) / np.abs(hist_1[:, :2]))
        print(f"  peer exchange: W {ePW:.2e}  H(block0) {ePH:.2e}  B {ePB:.2e}  loss rel err {p_err:.2e}  "
              f"W bit-identical across ranks: {same_w}")
        ok = ok and ePW < 1e-5 and ePH < 1e-5 and ePB < 1e-5 and p_err < 1e-5 and same_w
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print("dist_check OK")


if __name__ == "__main__":
    main()
