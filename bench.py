#!/usr/bin/env python
"""Benchmark of the B200-native MU-NMF loop (BASELINE.json metric: MU iterations/sec at 20,000 genes x 100,000
cells, k = 100, on 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # this framework (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores

One "step" is one full-batch MU iteration (W update, B updates, H update, loss terms).  Inputs are synthetic
(seeded, generated on the device), larger than L2 (X is 8 GB), so no L2 flush is needed between steps.

* ``value``    : iterations/s with X, Y, W, H, B resident in HBM; K timed steps between barrier+synchronize,
                 CUDA events on the launch stream, max over ranks.  At N > 1 the SAME 20k x 100k problem is
                 sharded by cells over the ranks ("scaling": "strong") with one NCCL all-reduce per iteration.
* ``e2e``      : the same metric through the public API (``ALPINE.fit`` on host numpy buffers whose rows are
                 page-locked): one fit of K iterations including validation, host->device upload of X / Y,
                 initialisation, the loop, the loss read-back and the device->host copy of W / H / B, divided by K;
                 after one untimed warm-up fit.
* ``roofline`` : the contraction kernel (two launches per step), timed with CUDA events around each launch
                 inside the timed region; HBM bytes of X against the measured HBM peak (tensor work: tf32-MMA equivalents against
                 measured bf16 peak / 2).
* ``cpu_baseline`` : the reference's step restated with its own torch-CPU operators (oracle/torch_port.py: its 7
                 GEMMs, randperm gather and G x n temporaries), measured on the full workload with all host cores
                 (a leading block of cells only when host memory or the time budget would be exceeded).
* ``gpu_torch_baseline`` : the same restatement on device="cuda" (fp32 cuBLAS, TF32 off): the reference's own GPU path
                 on this B200.
* ``cfg4`` / ``cfg5`` / ``parity_vs_n1`` : BASELINE configs[3] (CSR scaling workload) on the same ranks, configs[4]
                 (hyper-parameter search, at 8 GPUs) and the sharded == single-GPU self-check (N > 1).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = dict(
    name="cfg3: 20,000 genes x 100,000 cells dense fp32, k=100 (90 unguided + [5,5] guided, 3 and 4 categories), "
         "orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5, lam=[1e3,1e3], KL loss, full batch",
    n_genes=20000, n_cells=100000, n_components=90, n_covariate_components=[5, 5], categories=[3, 4],
    lam=[1e3, 1e3], orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5, eps=1e-6,
)
# BASELINE.json configs[3]: the sparse scaling workload (not the default bench line; `--workload cfg4`)
WORKLOAD_CFG4 = dict(
    name="cfg4: 30,000 genes x 1,000,000 cells CSR (5% density, 1 + Poisson(1) counts), k=100 (90 unguided + [5,5] "
         "guided, 3 and 4 categories), orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5, lam=[1e3,1e3], KL loss, full batch",
    n_genes=30000, n_cells=1000000, n_components=90, n_covariate_components=[5, 5], categories=[3, 4],
    lam=[1e3, 1e3], orth_W=0.2, alpha_W=0.5, l1_ratio_W=0.5, eps=1e-6, density=0.05,
)
METRIC = "MU iterations/sec at 20k genes x 100k cells, k=100"
UNIT = "iterations/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def load_traffic(key):
    """Measured DRAM bytes per launch of the contraction kernel (ncu --set full), profiles/r2_traffic.json."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f).get(key)
        vals = [v for v in (d.get("xh_bytes"), d.get("wx_bytes")) if v]
        return (sum(vals) / len(vals), d.get("source")) if vals else (None, None)
    except Exception:
        return None, None


def cublas_tf32_peak(dev, seconds=1.5):
    """cuBLAS TF32 GEMM throughput on this GPU (8192^3, fp32 in/out with allow_tf32): best of 10 launches (burst)
    and back to back for `seconds` (sustained, i.e. at the clocks the power cap allows).  MEASURED_PEAKS.json has no
    TF32 figure; SURVEY.md 8 d3 asks for one measured on the box."""
    import torch

    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn((n, n), device=dev)
        b = torch.randn((n, n), device=dev)
        c = torch.empty((n, n), device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize(dev)
        best = float("inf")
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(seconds * 1000.0 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize(dev)
        sustained = e0.elapsed_time(e1) / reps
        fl = 2.0 * n ** 3
        return {"burst": fl / best / 1e9, "sustained": fl / sustained / 1e9}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------- synthetic data
def synth_device_problem(dev, G, n_loc, col0, wl, seed=0):
    """Device-resident shard [col0, col0 + n_loc) of the synthetic problem: X (n_loc x G, cells-major), Ys, W, H, Bs.
    Low-rank + noise, non-negative.  Cells are generated in fixed GLOBAL chunks, each from its own seeded stream, so
    the data of a cell does not depend on how the cells are sharded: every N sees the same problem and the final
    loss must agree across N."""
    import torch

    from alpine_b200 import _native

    K = sum(wl["n_covariate_components"]) + wl["n_components"]
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + seed)
    rank = 16
    Wg = torch.rand((rank, G), device=dev, generator=g).pow_(3.0)
    X = _native.padded_rows(n_loc, G, dev)
    H = _native.padded_rows(K, n_loc, dev)
    Ys = [torch.empty((c, n_loc), dtype=torch.float32, device=dev) for c in wl["categories"]]
    step = max(1, (1 << 26) // max(G, 64))
    gs = torch.Generator(device=dev)
    for chunk in range(col0 // step, (col0 + n_loc + step - 1) // step if n_loc > 0 else 0):
        c0 = chunk * step
        gs.manual_seed(2000 + seed + 7919 * (chunk + 1))
        Hc = torch.rand((step, rank), device=dev, generator=gs).pow_(2.0)
        blk = Hc @ Wg
        blk.add_(torch.rand((step, G), device=dev, generator=gs).pow_(4.0), alpha=0.5)
        h0 = torch.rand((K, step), device=dev, generator=gs).clamp_(min=wl["eps"])
        codes = [torch.randint(0, c, (step,), device=dev, generator=gs) for c in wl["categories"]]
        a, b = max(c0, col0), min(c0 + step, col0 + n_loc)  # global cells of this chunk that belong to the shard
        X[a - col0:b - col0] = blk[a - c0:b - c0]
        H[:, a - col0:b - col0] = h0[:, a - c0:b - c0]
        for y, cd, c in zip(Ys, codes, wl["categories"]):
            y[:, a - col0:b - col0] = torch.nn.functional.one_hot(cd[a - c0:b - c0], c).T.float()
        del blk, Hc, h0
    gw = torch.Generator(device=dev)
    gw.manual_seed(42)
    W = torch.rand((G, K), device=dev, generator=gw).clamp_(min=wl["eps"])
    Bs = [torch.rand((c, k), device=dev, generator=gw).clamp_(min=wl["eps"]).contiguous()
          for c, k in zip(wl["categories"], wl["n_covariate_components"])]
    return X, Ys, W, H, Bs


def synth_device_csr(dev, G, n_loc, col0, wl, seed=0):
    """Device-resident CSR shard over cells: Bernoulli(density) mask x (1 + Poisson(1)) counts, built in fixed global
    chunks of cells (shard-independent, as synth_device_problem)."""
    import torch

    gs = torch.Generator(device=dev)
    step = max(1, (1 << 28) // G)
    counts, cols, vals = [], [], []
    for chunk in range(col0 // step, (col0 + n_loc + step - 1) // step if n_loc > 0 else 0):
        c0 = chunk * step
        gs.manual_seed(3000 + seed + 7919 * (chunk + 1))
        mask = torch.rand((step, G), device=dev, generator=gs) < wl["density"]
        nz = mask.nonzero()
        del mask
        v = 1.0 + torch.poisson(torch.ones(nz.shape[0], device=dev), generator=gs)
        a, b = max(c0, col0), min(c0 + step, col0 + n_loc)
        if a != c0 or b != c0 + step:  # keep the rows of this chunk that belong to the shard
            sel = (nz[:, 0] >= a - c0) & (nz[:, 0] < b - c0)
            nz, v = nz[sel], v[sel]
        counts.append(torch.bincount(nz[:, 0] - (a - c0), minlength=b - a))
        cols.append(nz[:, 1].to(torch.int32))
        vals.append(v)
        del nz
    indptr = torch.zeros(n_loc + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(torch.cat(counts), 0)
    return indptr, torch.cat(cols), torch.cat(vals).float()


# ------------------------------------------------------------------------------------------- CPU baseline
def _port_inputs(wl, n_cells, device="cpu", seed=0):
    """Synthetic inputs of the torch port (oracle/torch_port.py) on `device`: X as the genes x cells view of a
    cells-major buffer (main.py:104), labels, factors."""
    import torch

    from oracle import alpine_oracle as orc

    G = wl["n_genes"]
    g = torch.Generator(device=device).manual_seed(seed)
    Xcm = torch.empty((n_cells, G), device=device)
    step = max(1, (1 << 27) // G)
    for r0 in range(0, n_cells, step):  # chunked: no second 8 GB temporary next to X
        r1 = min(n_cells, r0 + step)
        Xcm[r0:r1] = torch.rand((r1 - r0, G), generator=g, device=device).pow_(3)
    Ys = []
    for c in wl["categories"]:
        codes = torch.randint(0, c, (n_cells,), generator=g, device=device)
        Ys.append(torch.nn.functional.one_hot(codes, c).T.contiguous().float())
    blocks = list(wl["n_covariate_components"]) + [wl["n_components"]]
    K = sum(blocks)
    W = torch.rand((G, K), generator=g, device=device).clamp_(min=1e-6)
    H = torch.rand((K, n_cells), generator=g, device=device).clamp_(min=1e-6)
    Bs = [torch.rand((c, k), generator=g, device=device).clamp_(min=1e-6)
          for c, k in zip(wl["categories"], wl["n_covariate_components"])]
    hp = orc.HyperParams(n_components=wl["n_components"], n_covariate_components=list(wl["n_covariate_components"]),
                         lam=list(wl["lam"]), orth_W=wl["orth_W"], alpha_W=wl["alpha_W"],
                         l1_ratio_W=wl["l1_ratio_W"], eps=wl["eps"])
    return Xcm.T, Ys, W, H, Bs, blocks, hp


def cpu_reference_run(wl, warmup, steps, budget_s=240.0):
    """The reference's iteration on the host CPU, MEASURED: oracle/torch_port.py restates main.py:502-663 (step, with
    its per-iteration randperm and X[:, perm] gather and its G x n temporaries) and main.py:726-753 (loss) with the
    same torch-CPU operators the reference runs on device="cpu" (MKL GEMMs, all host threads); pinned on the
    reference-generated fixtures.  Runs `warmup` + `steps` iterations on the FULL workload when host memory holds it
    (X + gather + two G x n temporaries: ~4.5 x |X|) and the run fits `budget_s`; otherwise on the largest leading
    block of cells that does, extrapolated linearly in cells (stated in "sample").  Returns a dict."""
    import psutil
    import torch

    from oracle import torch_port as tp

    threads = os.cpu_count() or 1
    old_threads = torch.get_num_threads()
    torch.set_num_threads(threads)  # torchrun exports OMP_NUM_THREADS=1
    try:
        G, n = wl["n_genes"], wl["n_cells"]
        avail = psutil.virtual_memory().available
        ns = n
        per_cell = 4.6 * 4.0 * G  # bytes of host memory per cell the port needs at its peak
        if per_cell * ns > 0.85 * avail:
            ns = max(1000, int(0.85 * avail / per_cell))
        # probe the cost per cell on a small block, then shrink the block if warmup + steps would overrun the budget
        X, Ys, W, H, Bs, blocks, hp = _port_inputs(wl, min(ns, 4000))
        tp.mu_step(X, Ys, W, H, Bs, blocks, hp, perm=torch.randperm(X.shape[1]))
        t0 = time.perf_counter()
        tp.mu_step(X, Ys, W, H, Bs, blocks, hp, perm=torch.randperm(X.shape[1]))
        tp.compute_loss(X, Ys, W, H, Bs, blocks, hp)
        per_cell_s = (time.perf_counter() - t0) / X.shape[1]
        del X, Ys, W, H, Bs
        total_iters = max(1, warmup) + max(1, steps)
        if per_cell_s * ns * total_iters > budget_s:
            ns = max(1000, int(budget_s / (per_cell_s * total_iters)))
        ns = min(ns, n)
        X, Ys, W, H, Bs, blocks, hp = _port_inputs(wl, ns)
        for _ in range(max(1, warmup)):
            tp.mu_step(X, Ys, W, H, Bs, blocks, hp, perm=torch.randperm(ns))
            tp.compute_loss(X, Ys, W, H, Bs, blocks, hp)
        t0 = time.perf_counter()
        for _ in range(max(1, steps)):
            tp.mu_step(X, Ys, W, H, Bs, blocks, hp, perm=torch.randperm(ns))  # main.py:502-663
            tp.compute_loss(X, Ys, W, H, Bs, blocks, hp)                      # main.py:666, 726-753
        dt = (time.perf_counter() - t0) / max(1, steps)
    finally:
        torch.set_num_threads(old_threads)
    full = ns == n
    value = 1.0 / (dt * (n / ns))
    sample = (f"the full workload ({n} cells x {G} genes), {max(1, warmup)} warm-up + {max(1, steps)} timed iterations, "
              f"{dt:.3f} s per iteration, nothing extrapolated" if full else
              f"the first {ns} of {n} cells (all {G} genes; host memory / time budget), {max(1, warmup)} warm-up + "
              f"{max(1, steps)} timed iterations, {dt:.3f} s per iteration on the sample, extrapolated linearly in cells")
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
            "seconds_per_iteration_measured": dt, "cells_measured": ns, "full_workload": full}


def gpu_torch_baseline(wl, dev, iters=3):
    """The reference's device="cuda" arithmetic on this B200 (the bar on the GPU): the same torch
    port as the CPU leg, on CUDA tensors, fp32 cuBLAS GEMMs with TF32 off (torch's default, the reference never
    switches it on), per-iteration randperm gather, G x n temporaries and the materialised-WH loss.  1 warm-up +
    `iters` timed iterations on the full workload, CUDA events."""
    import torch

    from oracle import torch_port as tp

    free, _ = torch.cuda.mem_get_info(dev)
    need = 5.5 * 4.0 * wl["n_genes"] * wl["n_cells"]
    if free < need:
        return {"value": None, "unit": UNIT, "skipped": f"needs {need / 2**30:.0f} GiB of free HBM"}
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        n = wl["n_cells"]
        X, Ys, W, H, Bs, blocks, hp = _port_inputs(wl, n, device=dev)

        def one():
            tp.mu_step(X, Ys, W, H, Bs, blocks, hp, perm=torch.randperm(n, device=dev))
            return tp.compute_loss(X, Ys, W, H, Bs, blocks, hp)

        one()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            one()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / iters
        peak = torch.cuda.max_memory_allocated(dev)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    del X, Ys, W, H, Bs
    torch.cuda.empty_cache()
    return {"value": 1000.0 / ms, "unit": UNIT, "ms_per_iteration": ms, "iterations": iters, "kind": "port",
            "what": "oracle/torch_port.py (the reference's torch operators, main.py:502-663 + 726-753) on device=cuda, "
                    "fp32 cuBLAS (allow_tf32=False), full workload, CUDA events",
            "peak_hbm_gb": peak / 1e9}


def run_reference_arm(args):
    """--impl reference: the reference's algorithm on the host CPU.  The reference itself is pure Python over
    torch and needs anndata/scanpy/kneed (absent, no network), and /root/reference does not exist on the GPU box, so
    the arm times the torch-CPU port of its iteration (oracle/torch_port.py, same operators and BLAS); rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOAD
    cb = cpu_reference_run(wl, args.warmup, args.steps)
    v = cb["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "l2": "inputs larger than L2"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
def measure_resident(args, wl, sparse, dev, world, rank, local_rank, steps, warmup, sample_clocks=True):
    """K timed full-batch iterations with everything resident in HBM (this rank's shard of `wl`); returns the fields
    of the JSON line that describe them (rank 0) or None."""
    import torch
    import torch.distributed as dist

    from alpine_b200 import _native
    from alpine_b200.engine import MUEngine, shard_bounds

    G, n = wl["n_genes"], wl["n_cells"]
    lo, hi = shard_bounds(n, world, rank)
    blocks = list(wl["n_covariate_components"]) + [wl["n_components"]]
    solver = _native.Solver(dev, G, hi - lo, blocks, wl["categories"])
    nnz = 0
    if sparse:
        wl_small = dict(wl, n_genes=4)  # labels and factors from the dense generator, X from the CSR generator
        _, Ys, _, H, Bs = synth_device_problem(dev, 4, hi - lo, lo, wl_small)
        gw = torch.Generator(device=dev)
        gw.manual_seed(42)
        W = torch.rand((G, sum(blocks)), device=dev, generator=gw).clamp_(min=wl["eps"])
        X = None
        csr = synth_device_csr(dev, G, hi - lo, lo, wl)
        nnz = int(csr[2].shape[0])
        solver.bind_csr(*csr)
        del csr
        torch.cuda.empty_cache()
    else:
        X, Ys, W, H, Bs = synth_device_problem(dev, G, hi - lo, lo, wl)
        solver.bind_dense(X)
    solver.bind_labels(Ys)
    solver.bind_factors(W, H, Bs)
    solver.set_hparams(wl["lam"], wl["alpha_W"], wl["l1_ratio_W"], wl["orth_W"], wl["eps"])
    exchange = "none (single GPU)"
    if world > 1:
        peer = (os.environ.get("ALPINE_B200_PEER", "1") != "0" and not args.use_als) and solver.enable_peer_exchange()
        exchange = ("NVLink peer memory inside the W-update kernels (reduce-scatter + update + all-gather)" if peer
                    else "NCCL all-reduce of the packed buffer")
    engine = MUEngine(solver, wl["lam"], use_als=args.use_als)
    # the K timed steps run without per-launch events (the events sit between the kernels and take the programmatic
    # dependent launch away from the contractions); a second pass of `steps_prof` steps right behind them, same
    # process and thermal state, carries the CUDA events that time every contraction launch for the roofline
    steps_prof = min(steps, 20)
    total = warmup + steps + steps_prof
    engine.begin(total)
    for it in range(warmup):
        engine.step(it)

    def fence():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    fence()
    sampler = ClockSampler(local_rank)
    if rank == 0 and sample_clocks:
        sampler.start()
    launches0 = _native.launch_count()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for it in range(warmup, warmup + steps):
        engine.step(it)
    e1.record()
    launches = _native.launch_count() - launches0
    solver.profile(True)
    for it in range(warmup + steps, total):
        engine.step(it)
    e2.record()
    fence()
    ms = e0.elapsed_time(e1)
    ms_prof = e1.elapsed_time(e2) / steps_prof
    solver.profile(False)
    gemm_ms, gemm_n = solver.profile_read()
    clocks = sampler.stop() if (rank == 0 and sample_clocks) else None
    t = torch.tensor([ms, gemm_ms / max(gemm_n, 1), ms_prof], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, gemm_ms_avg, ms_prof = float(t[0]), float(t[1]), float(t[2])
    hist = engine.collect_losses(total)  # also checks the kernels' error flag
    value = steps / (ms / 1000.0)

    # ---- roofline of the contraction kernel (per launch, this rank's shard)
    peaks = load_peaks()
    K = sum(blocks)
    n_loc = hi - lo
    # Split-precision product: one tf32 MMA (hi * hi) + two bf16 MMAs at twice the rate (hi * lo, lo * hi) = 2 tf32-MMA
    # equivalents of tensor time per fp32 product; tf32-exact integer counts (the CSR workload) need no lo * hi and keep
    # hi * lo as a tf32 MMA: also 2
    mma_per_product = 2.0
    flops = mma_per_product * 2.0 * G * n_loc * K
    tf32_peak = peaks["bf16_tflops_sustained"] / 2.0
    achieved = flops / (gemm_ms_avg * 1e-3) / 1e12 if gemm_ms_avg > 0 else 0.0
    x_bytes = 8.0 * nnz if sparse else 4.0 * G * n_loc
    hbm_ms = x_bytes / (peaks["hbm_gbs"] * 1e9) * 1e3
    tensor_ms = flops / (tf32_peak * 1e12) * 1e3
    achieved_gbs = x_bytes / (gemm_ms_avg * 1e-3) / 1e9 if gemm_ms_avg > 0 else 0.0
    traffic, traffic_src = (None, None)
    if world == 1 and not args.cells and not args.genes:
        traffic, traffic_src = load_traffic("cfg3_n1" if not sparse else "cfg4_n1")
    step_bound_ms = 2.0 * max(tensor_ms, hbm_ms)
    tensor_block = {"bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": achieved / tf32_peak, "bound_ms_per_launch": tensor_ms,
                    "what": "tf32-MMA equivalents: hi*hi as tf32 + the correction terms as bf16 MMAs at twice the rate"}
    # the same launch against SURVEY.md 8 d3's HBM bound (dense: 4 B per X element; CSR: 8 B per nonzero)
    hbm_block = {"bound": "hbm", "achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                 "frac": achieved_gbs / peaks["hbm_gbs"], "bound_ms_per_launch": hbm_ms}
    binding = hbm_block if hbm_ms >= tensor_ms else tensor_block  # the roofline that bounds this launch
    roofline = {
        "kernel": "mu_gemm_kernel (X H^T and W^T X, %s, tcgen05)" % (
            "2 tf32 MMAs per product on tf32-exact counts, CSR tile lists expanded on chip" if sparse
            else "tf32 hi*hi + bf16 hi*lo, lo*hi"),
        "bound": binding["bound"],
        "achieved": binding["achieved"], "peak": binding["peak"], "unit": binding["unit"], "frac": binding["frac"],
        "traffic": traffic, "traffic_source": traffic_src,
        "iteration": {"what": "whole MU iteration against the slower of (two contractions at the tensor peak, two "
                              "sweeps of X at HBM bandwidth) -- BASELINE north_star's roofline",
                      "bound_ms": step_bound_ms, "measured_ms": ms / steps,
                      "frac": step_bound_ms / (ms / steps)},
        "peak_source": f"{peaks['source']} hbm_gbs; tensor: bf16_tflops_sustained / 2 (no TF32 figure in "
                       f"MEASURED_PEAKS.json; the kernel is timed inside the step loop)",
        "avg_launch_ms": gemm_ms_avg, "launches_timed": gemm_n,
        # the pass the launches were timed in is a little slower per step than the timed region (an event pair sits
        # around every contraction launch): compare the kernel's share of the step within that pass
        "step_ms_in_timing_pass": ms_prof,
        "share_of_step": (2.0 * gemm_ms_avg / ms_prof) if ms_prof > 0 else None,
        "timed_in": f"{steps_prof} further steps right after the {steps} timed ones, with CUDA events around every "
                    f"contraction launch ({ms_prof:.3f} ms per step there)",
        "hbm": hbm_block, "tensor": tensor_block,
        # the roofline of the previous build (three tf32 MMAs per product, tensor bound), for comparison across rounds
        "vs_3xtf32_tensor_roofline": 3.0 * 2.0 * G * n_loc * K / (gemm_ms_avg * 1e-3) / 1e12 / tf32_peak
        if gemm_ms_avg > 0 and not sparse else None,
        "algorithmic": {"tf32_equiv_flop_per_launch": flops, "tf32_equiv_mma_per_fp32_product": mma_per_product,
                        "x_bytes_per_launch": x_bytes},
    }
    out = None
    if rank == 0:
        out = {
            "value": value, "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
            "config": {"workload": wl["name"] + (", use_als=True" if args.use_als else ""),
                       "parallelism": f"cells sharded over {world} GPU(s)", "exchange": exchange,
                       "l2": "inputs larger than L2 (X is %.1f GB per GPU)" % (x_bytes / 1e9)},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline,
            "final_loss": {"total": float(hist[-1, 0]), "reconstruction": float(hist[-1, 1])},
        }
    solver.close()
    del X, H, W, Ys, Bs, solver, engine
    torch.cuda.empty_cache()
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    sparse = args.workload == "cfg4"
    wl = dict(WORKLOAD_CFG4 if sparse else WORKLOAD)
    if args.cells:
        wl["n_cells"] = args.cells
    if args.genes:
        wl["n_genes"] = args.genes
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the MU loop has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)

    res = measure_resident(args, wl, sparse, dev, world, rank, local_rank, args.steps, args.warmup)
    line = None
    if rank == 0:
        line = {"metric": METRIC if not sparse else "MU iterations/sec at 30k genes x 1M cells CSR, k=100",
                "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 (split-precision tensor-core products: tf32 hi*hi + bf16 correction terms, fp32 accumulate)", "data": "synthetic",
                "config": res["config"], "clocks": res["clocks"], "gpu_launches": res["gpu_launches"],
                "roofline": res["roofline"], "final_loss": res["final_loss"]}
    if sparse:  # the sparse scaling workload has no host-API / CPU legs (the reference rejects sparse input)
        args.no_e2e = args.no_cpu = args.no_cfg4 = True

    # ---- e2e through the public API with host buffers (N = 1 process; each rank runs the sharded fit under torchrun)
    if not args.no_e2e:
        # every rank holds the full host matrix (the public API takes the same AnnData on every rank)
        import psutil

        need = 1.6 * world * 4.0 * wl["n_genes"] * wl["n_cells"]
        if psutil.virtual_memory().available < need:
            e2e = {"value": None, "unit": UNIT, "skipped": "not enough host memory for %d host copies of X" % world}
        else:
            e2e = run_e2e(args, wl, dev, world, rank)
        if rank == 0:
            line["e2e"] = e2e
    # ---- N > 1: sharded == single-GPU, once, outside every timed region (the test of tests/test_gpu_multi.py)
    if world > 1 and not args.no_parity:
        from alpine_b200.utils.dist_selfcheck import sharded_vs_single

        par = sharded_vs_single(dev)
        if rank == 0:
            line["parity_vs_n1"] = par
    # ---- BASELINE configs[3] on the same ranks: the workload the >= 6x @ 8 GPUs target is stated on
    if not args.no_cfg4 and not args.cells and not args.genes and not args.use_als:
        wl4 = dict(WORKLOAD_CFG4)
        free, _ = torch.cuda.mem_get_info(dev)
        # per GPU: tile lists 16 B/nnz + CSR source 12 B/nnz + generator scratch
        need4 = 30.0 * wl4["density"] * wl4["n_genes"] * (wl4["n_cells"] / world) + (6 << 30)
        if free < need4:
            c4 = {"skipped": "needs %.0f GiB of free HBM per GPU" % (need4 / 2**30)}
        else:
            r4 = measure_resident(args, wl4, True, dev, world, rank, local_rank, 30, 3, sample_clocks=False)
            c4 = None
            if rank == 0:
                c4 = {"metric": "MU iterations/sec at 30k genes x 1M cells CSR (5% density), k=100", "value": r4["value"],
                      "unit": UNIT, "ms_per_step": r4["ms_per_step"], "steps": 30, "warmup": 3, "scaling": "strong",
                      "config": r4["config"], "gpu_launches": r4["gpu_launches"], "final_loss": r4["final_loss"],
                      "roofline": {"tensor": r4["roofline"]["tensor"],
                                   "hbm": r4["roofline"]["hbm"], "avg_launch_ms": r4["roofline"]["avg_launch_ms"],
                                   "algorithmic": r4["roofline"]["algorithmic"]}}
        if rank == 0:
            line["cfg4"] = c4
    if rank == 0 and world == 1 and not args.no_minibatch and not args.cells and not args.genes:
        try:
            line["minibatch"] = run_minibatch_block(dev)
        except Exception as exc:  # never lose the headline line to an auxiliary block
            line["minibatch"] = {"error": repr(exc)[:300]}
    if rank == 0 and world == 1 and not args.no_tf32_peak:
        # after every timed region, so that its heat does not touch them
        tf = cublas_tf32_peak(dev)
        line["roofline"]["cublas_tf32_tflops"] = tf
        line["roofline"]["tensor"]["frac_of_cublas_tf32_sustained"] = line["roofline"]["tensor"]["achieved"] / tf["sustained"]
    if rank == 0 and world == 1 and not args.no_gpu_torch:
        line["gpu_torch_baseline"] = gpu_torch_baseline(wl, dev)
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_reference_run(wl, warmup=1, steps=2, budget_s=40.0)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    # ---- BASELINE configs[4]: 64 trials x 3 folds, one fit per GPU, from ONE process (rank 0) over all GPUs of the box
    if rank == 0 and (args.cfg5 or (world == 8 and not args.no_cfg5)):
        try:
            line["cfg5"] = run_cfg5(args)
        except Exception as exc:  # never lose the headline line to the auxiliary block
            line["cfg5"] = {"error": repr(exc)[:300]}
    if rank == 0:
        print(json.dumps(line), flush=True)


def run_minibatch_block(dev, batch_size=4096):
    """Mini-batch epochs (main.py:509-521) at BASELINE configs[1] shapes (5,000 HVG x 50,000 cells, 30 + [5, 5]) through
    the public API: epochs/s from the difference of a 22-epoch and a 2-epoch fit, the faster of three runs each (upload,
    init and download cancel; a 4-epoch difference was within the host-side jitter of one fit).
    One epoch = 13 batches of 4,096 cells (gather, one MU step on the batch, scatter) + the full-data loss."""
    import pandas as pd
    import torch

    from alpine_b200 import ALPINE
    from alpine_b200.utils.anndata_compat import AnnData
    from alpine_b200.utils.synth import make_labels

    n, G = 50_000, 5_000
    g = torch.Generator(device=dev).manual_seed(9)
    X = (torch.rand((n, 12), device=dev, generator=g).pow_(2.0) @ torch.rand((12, G), device=dev, generator=g).pow_(3.0)
         + 0.25 * torch.rand((n, G), device=dev, generator=g).pow_(4.0)).cpu().numpy()
    labels = make_labels(n, [3, 4], seed=9)
    obs = pd.DataFrame({f"cov{i}": pd.Series(l, dtype=object) for i, l in enumerate(labels)})
    kw = dict(n_components=30, n_covariate_components=[5, 5], lam=[1e3, 1e3], device=str(dev))
    times = {}
    short, long_ = 2, 22
    for epochs in (short, short, long_, short, long_, short, long_):  # the first fit warms up
        model = ALPINE(**kw)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        model.fit(AnnData(X, obs=obs.copy()), ["cov0", "cov1"], batch_size=batch_size, max_iter=epochs)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        times[epochs] = min(times.get(epochs, dt), dt)
    per_epoch = (times[long_] - times[short]) / float(long_ - short)
    return {"what": f"ALPINE.fit(batch_size={batch_size}) on {G} genes x {n} cells, 30 + [5, 5] components, random sampler",
            "epochs_per_s": 1.0 / per_epoch, "ms_per_epoch": 1000.0 * per_epoch, "batches_per_epoch": -(-n // batch_size),
            "fit_seconds": {str(k): round(v, 4) for k, v in times.items()},
            "loss_total_last": float(model.loss_history["total loss"].iloc[-1])}


def run_cfg5(args, trials=64, n_splits=3, max_iter=100):
    """ComponentOptimizer.search_hyperparams on 5,000 HVG x 50,000 cells, `trials` trials x `n_splits` folds, one fit
    per GPU over every visible GPU (worker threads of this process; the other ranks have exited)."""
    import pandas as pd
    import torch

    from alpine_b200.optimization import ComponentOptimizer
    from alpine_b200.utils.anndata_compat import AnnData
    from alpine_b200.utils.synth import make_labels

    n, G = 50_000, 5_000
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    Wg = torch.rand((12, G), device=dev, generator=g).pow_(3.0)
    Hc = torch.rand((n, 12), device=dev, generator=g).pow_(2.0)
    labels = make_labels(n, [3, 4], seed=5)
    # cells of the same joint label share a bump in two latent factors, so that the embedding can be scored
    codes = torch.from_numpy(np.stack([pd.factorize(l)[0] for l in labels], 1)).to(dev)
    Hc[torch.arange(n, device=dev), codes[:, 0]] += 1.0
    Hc[torch.arange(n, device=dev), 3 + codes[:, 1]] += 1.0
    X = (Hc @ Wg + 0.25 * torch.rand((n, G), device=dev, generator=g).pow_(4.0)).cpu().numpy()
    del Wg, Hc
    torch.cuda.empty_cache()
    obs = pd.DataFrame({f"cov{i}": pd.Series(l, dtype=object) for i, l in enumerate(labels)})
    opt = ComponentOptimizer(AnnData(X, obs=obs), ["cov0", "cov1"], max_iter=max_iter, device="cuda")
    t0 = time.perf_counter()
    best = opt.search_hyperparams(n_total_components_range=(10, 100), max_evals=trials, n_splits=n_splits)
    wall = time.perf_counter() - t0
    busy = {d: round(t / wall, 3) for d, t in sorted(opt.device_busy_s.items())}
    return {"what": f"ComponentOptimizer.search_hyperparams, {trials} trials x {n_splits} folds on {G} HVG x {n} cells, "
                    f"max_iter={max_iter}, one fit per GPU (fit on 2/3 of the cells + transform of 1/3 + scoring per fold)",
            "wall_s": wall, "s_per_trial": wall / trials, "gpus": len(opt.devices), "trials": trials,
            "per_gpu_busy_fraction": busy,
            "optimiser": "seeded random search stand-in (hyperopt is not installed); k-means scorer stand-in (scanpy absent)",
            "best": {k: (v if not isinstance(v, float) else round(v, 4)) for k, v in best.items()}}


def run_e2e(args, wl, dev, world, rank):
    """ALPINE.fit on host numpy data: upload, init, K iterations, loss read-back, factors back to the host."""
    import pandas as pd
    import torch

    from alpine_b200 import ALPINE
    from alpine_b200.utils.anndata_compat import AnnData

    G, n = wl["n_genes"], wl["n_cells"]
    from alpine_b200.engine import shard_bounds

    # host data (built on the device for speed, then copied out; not timed).  The rows this rank uploads are
    # page-locked (cudaHostRegister), as the bench contract asks ("from pinned host memory"): the library's uploader
    # then hands them to the DMA engine directly; pageable arrays go through its staging ring instead
    # (tools/e2e_probe.py measures both).
    X, Ys, _, _, _ = synth_device_problem(dev, G, n, 0, wl, seed=1)
    Xh = X.cpu().numpy()
    lo, hi = shard_bounds(n, world, rank)
    pinned = False
    if not args.e2e_pageable and hi > lo:
        rc = torch.cuda.cudart().cudaHostRegister(Xh.ctypes.data + lo * G * 4, (hi - lo) * G * 4, 0)
        pinned = int(rc) == 0
    obs = {}
    for i, y in enumerate(Ys):
        codes = y.argmax(dim=0).cpu().numpy()
        obs[f"cov{i}"] = pd.Series([f"c{v}" for v in codes], dtype=object)
    del X, Ys
    torch.cuda.empty_cache()
    adata = AnnData(Xh, obs=pd.DataFrame(obs))
    model = ALPINE(n_components=wl["n_components"], n_covariate_components=list(wl["n_covariate_components"]),
                   lam=list(wl["lam"]), orth_W=wl["orth_W"], alpha_W=wl["alpha_W"], l1_ratio_W=wl["l1_ratio_W"],
                   device=str(dev))
    keys = list(obs.keys())
    steps = args.steps
    if world > 1:
        import torch.distributed as dist

        # NCCL opens its channels lazily on the first collective of each size class (0.3-0.8 s, once per process,
        # like CUDA context creation); the device-resident arm above exchanged over peer memory, so do that here,
        # outside the timed fit
        for numel in (1 << 21, 1 << 23):
            warm = torch.zeros(numel, device=dev)
            dist.all_reduce(warm)
        del warm
        dist.barrier()
    # One untimed fit first, as the W warm-up steps of the resident arm: the first fit of a process pays for lazily
    # loaded CUDA modules and for cudaMalloc calls that torch's caching allocator then keeps (the device-resident arm
    # above emptied the cache).  Its time is reported next to the timed fit.
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    model.fit(adata, keys, max_iter=steps)
    torch.cuda.synchronize(dev)
    first_fit = time.perf_counter() - t0
    adata = AnnData(Xh, obs=pd.DataFrame(obs))
    model = ALPINE(n_components=wl["n_components"], n_covariate_components=list(wl["n_covariate_components"]),
                   lam=list(wl["lam"]), orth_W=wl["orth_W"], alpha_W=wl["alpha_W"], l1_ratio_W=wl["l1_ratio_W"],
                   device=str(dev))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    model.fit(adata, keys, max_iter=steps)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt[0])
    K = model.total_components
    n_cat = sum(wl["categories"])
    h2d = 4.0 * (G * n + n_cat * n) / world
    d2h = 4.0 * (G * K + K * n + sum(c * k for c, k in zip(wl["categories"], wl["n_covariate_components"]))) + 8.0 * steps * 4
    peer_fit = os.environ.get("ALPINE_B200_PEER", "0") == "1"
    if pinned:
        torch.cuda.cudart().cudaHostUnregister(Xh.ctypes.data + lo * G * 4)
    return {"value": steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d / steps, "d2h_bytes_per_step": d2h / steps,
            "host_memory": "page-locked (cudaHostRegister on this rank's rows of adata.X)" if pinned else "pageable",
            "exchange": ("none (single GPU)" if world == 1 else
                         "NVLink peer memory (ALPINE_B200_PEER=1)" if peer_fit else
                         "NCCL all-reduce (ALPINE.fit's default: mapping peer memory costs ~0.2 s per fit)"),
            "seconds_per_fit": dt, "iterations_per_fit": steps, "first_fit_seconds_untimed_warmup": first_fit, "nccl_channels_warmed_before_timing": world > 1,
            "phases_s": {k: round(v, 4) for k, v in getattr(model, "timings", {}).items()},
            "phases_detail_s": {k: round(v, 4) for k, v in getattr(model, "timings_detail", {}).items()},
            "what": "ALPINE(...).fit(adata, keys, max_iter=steps) on host numpy data: validation, encoders, H2D of X/Y, "
                    "init, the loop, loss read-back, scaling, D2H of W/H/B, store_embeddings"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=0, help="override the workload's cell count (debugging)")
    ap.add_argument("--genes", type=int, default=0, help="override the workload's gene count (debugging)")
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg4"],
                    help="cfg3 = the headline dense workload; cfg4 = BASELINE configs[3], 30k x 1M CSR (scaling study)")
    ap.add_argument("--use-als", action="store_true",
                    help="time the block Gauss-Seidel sweep (use_als=True, main.py:523-588) instead of the default update")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-pageable", action="store_true", help="leave adata.X in pageable host memory for the e2e fit")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-tf32-peak", action="store_true", help="skip the cuBLAS TF32 reference measurement")
    ap.add_argument("--no-gpu-torch", action="store_true", help="skip the torch-CUDA (reference arithmetic) baseline leg")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the auxiliary cfg4 (CSR scaling workload) block")
    ap.add_argument("--no-minibatch", action="store_true", help="skip the auxiliary mini-batch block (cfg2 shapes)")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the auxiliary cfg5 (hyper-parameter search) block at 8 GPUs")
    ap.add_argument("--cfg5", action="store_true", help="run the cfg5 block at any GPU count")
    ap.add_argument("--no-parity", action="store_true", help="skip the sharded == single-GPU self-check at N > 1")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
