"""Numerics of the contraction kernel's split-precision product, emulated on the CPU.

The CUDA kernel (csrc/mu_gemm_sm100.cuh) computes x * y as  tf32(x) * tf32(y)  [tf32 MMA]
+ bf16(hi_x) * bf16(lo_y) + bf16(lo_x) * bf16(hi_y)  [bf16 MMAs],  hi = tf32(x), lo = x - hi.  This test runs the
oracle's MU loop on reference-generated goldens with EVERY matrix product replaced by that scheme (operand roundings
exactly as in csrc/ptx_sm100.cuh, products and sums in fp64) and checks that the trajectory stays as close to the
reference's as with plain 3xTF32 (all three terms in tf32) -- and that dropping the correction terms does not, i.e. that
the check is sensitive.  It documents why the cheaper correction terms are admissible; the kernel itself is tested
against the same goldens in tests/test_gpu_parity.py.
"""
from __future__ import annotations

import numpy as np
import pytest

from oracle import alpine_oracle as orc
from tests.helpers import epoch_batches, hp_of, inputs_of, load_golden, rel_fro


def _tf32(x):  # ptx::round_tf32: round to nearest (ties away) on the 13 dropped mantissa bits
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def _bf16(x):  # cvt.rn.bf16.f32: round to nearest even
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = b + np.uint32(0x7FFF) + ((b >> np.uint32(16)) & np.uint32(1))
    return (r & np.uint32(0xFFFF0000)).view(np.float32)


def _product(a, b, mode):
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    ah, bh = _tf32(a), _tf32(b)
    al, bl = a - ah, b - bh  # exact in fp32
    d = np.float64
    if mode == "3xtf32":
        p = ah.astype(d) @ bh.astype(d) + ah.astype(d) @ _tf32(bl).astype(d) + _tf32(al).astype(d) @ bh.astype(d)
    elif mode == "tf32+bf16":
        p = (ah.astype(d) @ bh.astype(d) + _bf16(ah).astype(d) @ _bf16(bl).astype(d)
             + _bf16(al).astype(d) @ _bf16(bh).astype(d))
    elif mode == "1xtf32":
        p = ah.astype(d) @ bh.astype(d)
    else:
        raise ValueError(mode)
    return p.astype(np.float32)


def _run(name, mode):
    class Emu(np.ndarray):  # every `@` of the oracle goes through the emulated product
        def __matmul__(self, other):
            return _product(self, other, mode).view(Emu)

        def __rmatmul__(self, other):
            return _product(other, self, mode).view(Emu)

    g = load_golden(name)
    hp = hp_of(name)
    X, Ys, st = inputs_of(g)
    X, Ys = X.view(Emu), [y.view(Emu) for y in Ys]
    st.W, st.H, st.Bs = st.W.view(Emu), st.H.view(Emu), [b.view(Emu) for b in st.Bs]
    last = int(max(g["kept_iters"]))
    for it in range(1, last + 1):
        for idx in epoch_batches(g, it):
            orc.mu_step(X, Ys, st, hp, idx=idx)
    return max(rel_fro(np.asarray(st.W), g[f"W_it{last}"]), rel_fro(np.asarray(st.H), g[f"H_it{last}"]))


@pytest.mark.parametrize("name", ["kl_basic", "kl_reg_nan", "frob_reg"])
def test_bf16_correction_terms_track_the_reference_like_3xtf32(name):
    dev3 = _run(name, "3xtf32")
    dev2 = _run(name, "tf32+bf16")
    dev1 = _run(name, "1xtf32")
    assert dev3 < 2e-6 and dev2 < 2e-6, (dev3, dev2)   # both at the level of fp32 summation-order noise
    assert dev2 < 2.5 * dev3, (dev3, dev2)
    assert dev1 > 50 * dev2, (dev1, dev2)              # without correction terms: tf32-level error, 100x larger


@pytest.mark.parametrize("name", ["kl_long200", "kl_scores2k"])
def test_bf16_correction_terms_do_not_drift_over_long_trajectories(name):
    """200 iterations on 500 x 300 and 60 iterations on 2,000 genes: where rounding noise has the time to be amplified
    by the updates, both schemes end equally far (1-2e-6) from the reference's trajectory."""
    dev3 = _run(name, "3xtf32")
    dev2 = _run(name, "tf32+bf16")
    assert dev3 < 3e-6 and dev2 < 3e-6, (dev3, dev2)
    assert dev2 < 2.0 * dev3, (dev3, dev2)

