"""Host logic of the contraction kernel's work split (stream-K inside L2-sized pieces): the segment numbering the
kernel follows must agree with the slot lists the reduce kernels are given.  Compiled for the host with nvcc (no GPU)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "csrc", "worksplit_check.cu")
EXE = os.path.join(ROOT, "tests", "csrc", "worksplit_check.bin")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


@pytest.fixture(scope="module")
def exe():
    if not os.path.isfile(NVCC):
        pytest.skip("nvcc not available")
    deps = [SRC, os.path.join(ROOT, "alpine_b200", "csrc", "mu_gemm_sm100.cuh")]
    if not os.path.isfile(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-o", EXE, SRC], check=True)
    return EXE


CASES = [
    # num_tiles, kb_per_tile, piece_len, grid
    (79, 3125, 1048, 148),    # cfg3 X H^T  (20k genes, 100k cells)
    (391, 625, 625, 148),     # cfg3 W^T X
    (118, 31250, 984, 148),   # cfg4 X H^T on one GPU (1M cells)
    (1, 3125, 3125, 148),     # Gram H H^T
    (1, 5, 5, 5),             # fewer units than SMs
    (3, 17, 8, 7),
    (2, 1, 1, 2),
    (20, 157, 40, 148),
    (5, 4, 4, 148),           # more CTAs than units in a piece
]


@pytest.mark.parametrize("case", CASES)
def test_segments_match_reduce_slot_lists(exe, case):
    r = subprocess.run([exe] + [str(v) for v in case], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr


def test_random_work_spaces(exe):
    """Seeded random (tiles, k-blocks, piece length, grid) combinations, including grids larger than the work."""
    import random

    rng = random.Random(1234)
    for _ in range(60):
        tiles = rng.choice([1, 2, 3, 7, 40, 118, 391, 1000])
        kb = rng.choice([1, 2, 5, 33, 157, 625, 3125, 9000])
        piece = max(1, min(kb, rng.choice([1, 4, 8, 64, 312, 1048, 100000])))
        grid = rng.choice([1, 2, 5, 64, 132, 148])
        r = subprocess.run([exe, str(tiles), str(kb), str(piece), str(grid)], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.startswith("OK"), (tiles, kb, piece, grid, r.stdout + r.stderr)
