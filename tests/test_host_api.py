"""CPU tests of the host side: C-ABI surface, drop-in API validation, encoders, elbow detection, no-CPU-fallback."""
import os
import re

import numpy as np
import pandas as pd
import pytest

from alpine_b200 import _native
from alpine_b200.main import ALPINE
from alpine_b200.utils.anndata_compat import AnnData
from alpine_b200.utils.encoder import FeatureEncoders
from alpine_b200.utils.kneedle import find_elbow
from oracle import alpine_oracle as orc
from tests.helpers import golden_names, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "alpine_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(alpine_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    """The C-ABI library loads (no GPU needed for dlopen) and exports everything include/alpine_b200.h declares."""
    import __graft_entry__ as ge

    ge.build()
    lib = _native.load_library()
    names = _header_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name
        assert name in _native.SIGNATURES, f"{name} has no ctypes prototype"
    assert sorted(_native.SIGNATURES) == names
    assert lib.alpine_abi_version() >= 2
    assert lib.alpine_launch_count() == 0  # nothing launched on a machine without a GPU


def test_create_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_native.AlpineNativeError):
        _native.Solver("cuda:0", 10, 10, [4], [])


def _toy_adata(n=40, G=12, seed=0, with_nan=False):
    rng = np.random.default_rng(seed)
    X = rng.gamma(0.3, 2.0, size=(n, G)).astype(np.float32)
    lab = np.array([f"b{v}" for v in rng.integers(0, 3, n)], dtype=object)
    if with_nan:
        lab[::7] = np.nan
    obs = pd.DataFrame({"batch": pd.Series(lab, dtype=object), "num": np.arange(n)})
    obs.index = [str(i) for i in range(n)]
    return AnnData(X, obs=obs)


def test_constructor_validation_matches_reference_behaviour():
    ok = dict(n_components=5, n_covariate_components=[2], lam=[1.0])
    m = ALPINE(**ok)
    assert m.n_all_components == [2, 5] and m.total_components == 7
    assert ALPINE(l1_ratio=0.25, **ok).l1_ratio_W == 0.25  # north_star alias
    bad = [
        (dict(ok, n_components=0), ValueError),
        (dict(ok, n_covariate_components=(2,)), TypeError),
        (dict(ok, n_covariate_components=[-1]), ValueError),
        (dict(ok, lam=(1.0,)), TypeError),
        (dict(ok, lam=[1]), ValueError),            # ints are rejected: main.py:342
        (dict(ok, alpha_W=1), ValueError),          # main.py:348
        (dict(ok, orth_W=-0.1), ValueError),
        (dict(ok, l1_ratio_W=1.5), ValueError),
        (dict(ok, scale_needed=1), TypeError),
        (dict(ok, loss_type="l2"), ValueError),
        (dict(ok, loss_type=3), TypeError),
        (dict(ok, eps=0), ValueError),
        (dict(ok, random_state=-1), ValueError),
    ]
    for kw, exc in bad:
        with pytest.raises(exc):
            ALPINE(**kw)


def test_fit_validation_and_no_cpu_fallback():
    ad = _toy_adata()
    m = ALPINE(n_components=4, n_covariate_components=[2], lam=[10.0], device="cpu")
    with pytest.raises(TypeError):
        m.fit("not adata", ["batch"])
    with pytest.raises(TypeError):
        m.fit(ad, "batch")
    with pytest.raises(ValueError):
        m.fit(ad, ["batch", "other"])
    with pytest.raises(ValueError):
        m.fit(ad, ["missing"])
    with pytest.raises(TypeError):
        m.fit(ad, ["num"])  # not an object column: main.py:415
    neg = _toy_adata()
    neg.X[0, 0] = -1
    with pytest.raises(ValueError):
        m.fit(neg, ["batch"])
    with pytest.raises(RuntimeError):
        m.transform(ad)  # not fitted
    with pytest.raises(RuntimeError):
        m.get_covariate_gene_scores()
    # a CPU device is refused: the MU loop has no CPU path
    with pytest.raises(_native.AlpineNativeError):
        m.fit(ad, ["batch"], max_iter=2)


@pytest.mark.parametrize("name", golden_names())
def test_feature_encoders_match_reference_fixture(name):
    g = load_golden(name)
    n_cov = int(g["n_cov"])
    keys = [f"cov{i}" for i in range(n_cov)]
    cols = {}
    for i, k in enumerate(keys):
        lab = np.array([np.nan if na else str(v) for v, na in zip(g[f"labels{i}"], g[f"labels{i}_isna"])], dtype=object)
        cols[k] = pd.Series(lab, dtype=object)
    df = pd.DataFrame(cols)
    fe = FeatureEncoders(keys)
    Y = fe.fit_transform(df)
    for i, k in enumerate(keys):
        np.testing.assert_array_equal(Y[i], g[f"Y{i}_cells_by_cat"])
        assert fe.encoded_labels[k] == [str(c) for c in g[f"cats{i}"]]
        np.testing.assert_array_equal(fe.transform(df)[i], Y[i])
    # unseen category at transform time -> all-zero row (handle_unknown="ignore")
    if n_cov:
        df2 = df.copy()
        df2.loc[0, keys[0]] = "never_seen"
        assert fe.transform(df2)[0][0].sum() == 0


def test_elbow_matches_oracle_restatement():
    x = np.arange(200)
    for rate, floor in ((0.05, 1.0), (0.02, 3.0), (0.2, 0.5)):
        y = np.log10(100.0 * np.exp(-rate * x) + floor)
        assert find_elbow(x, y) == orc.kneedle_elbow(x, y)
    assert find_elbow(x, np.log10(100.0 * np.exp(-0.05 * x) + 1.0)) is not None
    assert find_elbow(np.arange(2), np.array([1.0, 0.5])) is None


def test_anndata_standin_subsetting_and_copy():
    ad = _toy_adata(with_nan=True)
    ad.obsm["e"] = np.arange(ad.shape[0] * 2).reshape(-1, 2)
    sub = ad[np.array([3, 1, 5])]
    assert sub.shape == (3, ad.shape[1])
    np.testing.assert_array_equal(sub.X, ad.X[[3, 1, 5]])
    assert list(sub.obs.index) == ["3", "1", "5"]
    np.testing.assert_array_equal(sub.obsm["e"], ad.obsm["e"][[3, 1, 5]])
    cp = ad.copy()
    cp.X[0, 0] = 123.0
    assert ad.X[0, 0] != 123.0
    assert ad.var_names.tolist() == [str(i) for i in range(ad.shape[1])]


def test_csr_row_gather_matches_scipy():
    """Device-side CSR row gather of the sparse mini-batch path (runs on CPU tensors here)."""
    import scipy.sparse as sp
    import torch

    from alpine_b200.main import _csr_take_rows

    rng = np.random.default_rng(0)
    M = sp.random(60, 37, density=0.15, format="csr", random_state=2, dtype=np.float32)
    M.data[:] = rng.random(M.nnz).astype(np.float32) + 0.1
    M[5] = 0  # an empty row
    M.eliminate_zeros()
    csr = (torch.from_numpy(M.indptr.astype(np.int64)), torch.from_numpy(M.indices.astype(np.int32)),
           torch.from_numpy(M.data.copy()))
    for idx in (rng.permutation(60)[:23], rng.integers(0, 60, 50), np.array([5]), np.array([5, 5, 7])):
        ip, ii, vv = _csr_take_rows(csr, torch.from_numpy(idx.astype(np.int64)))
        got = sp.csr_matrix((vv.numpy(), ii.numpy(), ip.numpy()), shape=(len(idx), 37)).toarray()
        np.testing.assert_array_equal(got, M[idx].toarray())


def test_result_buffers_give_the_arrays_the_reference_stores():
    """fit() pre-allocates its result arrays on a background thread (main._HostBuffers); store_embeddings fills them
    instead of calling copy(): same values, same memory order as the reference's copy(H.T) / copy(W) (main.py:303-320),
    independent of the model's own matrices."""
    from copy import copy

    from alpine_b200.main import _HostBuffers

    rng = np.random.default_rng(0)
    G, n, ks = 7, 11, [2, 3]
    W = rng.random((G, sum(ks)), dtype=np.float32)
    H = rng.random((sum(ks), n), dtype=np.float32)
    model = ALPINE(n_components=3, n_covariate_components=[2], lam=[1.0])
    model.covariate_keys = ["c"]
    model.matrices = {"Ws": [np.ascontiguousarray(W[:, :2]), np.ascontiguousarray(W[:, 2:])], "Hs": [H[:2], H[2:]]}
    obs = pd.DataFrame({"c": ["a", "b"] * 5 + ["a"]})
    Y = [np.zeros((n, 2), dtype=np.float32)]

    plain = AnnData(np.zeros((n, G), dtype=np.float32), obs=obs.copy())
    model.store_embeddings(plain, _dummy_matrices=Y)
    bufs = _HostBuffers({f"{kind}{i}": shape for i, k in enumerate(ks)
                         for kind, shape in (("obsm", (k, n)), ("varm", (G, k)))})
    fast = AnnData(np.zeros((n, G), dtype=np.float32), obs=obs.copy())
    model.store_embeddings(fast, _dummy_matrices=Y, _bufs=bufs)
    for slot in ("ALPINE_embedding", "c"):
        a, b = plain.obsm[slot], fast.obsm[slot]
        assert a.shape == b.shape and a.strides == b.strides and np.array_equal(a, b)
        assert np.array_equal(a, copy(model.matrices["Hs"][-1 if slot == "ALPINE_embedding" else 0].T))
        assert not np.shares_memory(b, H)
    for slot in ("ALPINE_weights", "c"):
        a, b = plain.varm[slot], fast.varm[slot]
        assert a.shape == b.shape and a.strides == b.strides and np.array_equal(a, b)
    assert bufs.take("obsm0", (2, n)) is None          # handed out once
    assert _HostBuffers({"x": (3, 4)}).take("x", (4, 3)) is None  # a shape that was not prepared: caller allocates
