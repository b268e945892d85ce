"""Run under torchrun with N >= 2 GPUs: cell-sharded fit (NCCL all-reduce and NVLink peer exchange) == single-GPU fit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py

Prints one JSON line on rank 0 and exits non-zero when a bar is missed (alpine_b200/utils/dist_selfcheck.py).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpine_b200.utils.dist_selfcheck import minibatch_sharded_vs_single, sharded_vs_single  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=6001)
    ap.add_argument("--genes", type=int, default=1500)
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--blocks", default="5,5,90")
    ap.add_argument("--cats", default="3,4")
    ap.add_argument("--minibatch", action="store_true", help="check mini-batch epochs under cell sharding instead")
    args = ap.parse_args()
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    if args.minibatch:
        parts = [minibatch_sharded_vs_single(dev, sampling_method=m, use_als=a)
                 for m, a in (("random", False), ("weighted", False), ("random", True))]
        res = {"world": parts[0]["world"], "cases": parts, "ok": all(p["ok"] for p in parts)}
    else:
        res = sharded_vs_single(dev, n=args.cells, G=args.genes, blocks=[int(v) for v in args.blocks.split(",")],
                                cats=[int(v) for v in args.cats.split(",")], n_iter=args.iters)
    if rank == 0:
        print("dist_check " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()
