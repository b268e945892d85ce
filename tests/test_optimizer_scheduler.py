"""CPU tests of the trial scheduler ("one fit per GPU") and the ComponentOptimizer shell, with a stand-in model."""
import threading
import time

import numpy as np
import pandas as pd
import pytest

from alpine_b200.optimization import ComponentOptimizer, STATUS_FAIL, STATUS_OK
from alpine_b200.scheduler import DeviceScheduler
from alpine_b200.utils.anndata_compat import AnnData


def test_scheduler_runs_one_job_per_device_at_a_time():
    devices = ["cuda:0", "cuda:1", "cuda:2"]
    active, peak, lock = {d: 0 for d in devices}, {d: 0 for d in devices}, threading.Lock()

    def job(x, device):
        with lock:
            active[device] += 1
            peak[device] = max(peak[device], active[device])
        time.sleep(0.02)
        with lock:
            active[device] -= 1
        return x * x

    sch = DeviceScheduler(devices)
    out = sch.map(job, list(range(10)))
    assert out == [i * i for i in range(10)]           # results in job order
    assert all(v == 1 for v in peak.values())           # never two jobs on one device
    assert {d for _, d in sch.assignments} == set(devices)
    assert sorted(i for i, _ in sch.assignments) == list(range(10))
    with pytest.raises(ZeroDivisionError):
        DeviceScheduler(devices).map(lambda x, d: 1 // x, [1, 0, 2])
    assert DeviceScheduler(["cpu"]).map(lambda x, d: (x, d), [1, 2]) == [(1, "cpu"), (2, "cpu")]


class _FakeModel:
    """Stands in for ALPINE: records the device, writes a label-informative embedding."""

    seen = []

    def __init__(self, **kw):
        self.kw = kw
        self.max_iter = 7

    def fit(self, adata, covariate_keys, max_iter=None, **_):
        _FakeModel.seen.append(self.kw["device"])
        self.keys = covariate_keys
        return self

    def store_embeddings(self, adata):
        pass

    def transform(self, adata):
        codes = pd.Categorical(adata.obs[self.keys[0]]).codes.astype(np.float32)
        rng = np.random.default_rng(0)
        adata.obsm["ALPINE_embedding"] = np.stack([codes, codes * 0.5], axis=1) + 0.01 * rng.normal(size=(len(codes), 2))


def _adata(n=90, G=8):
    rng = np.random.default_rng(0)
    obs = pd.DataFrame({"batch": pd.Series([f"b{i % 3}" for i in range(n)], dtype=object),
                        "cond": pd.Series([f"c{i % 2}" for i in range(n)], dtype=object)})
    return AnnData(rng.random((n, G), dtype=np.float32), obs=obs)


def test_component_split_rule_and_objective_status():
    opt = ComponentOptimizer(_adata(), ["batch", "cond"], max_iter=5, device="cpu")
    opt.min_covariate_components = [3, 2]
    n, guided = opt._distribute_components({"n_total_components": 40.0, "splits": [0.5, 0.25, 0.25]})
    assert guided == [10, 5] and n == 25            # rest = 20 split 2:1, unguided gets the remainder
    n, guided = opt._distribute_components({"n_total_components": 10.0, "splits": [0.01, 0.01, 0.98]})
    assert guided == [3, 2] and n == 5                # floored at min_covariate_components
    opt.n_splits, opt.iter_records = 2, []
    bad = opt.objective({"n_total_components": 6.0, "splits": [0.5, 0.5, 0.0001], "lam_0": 1.0, "lam_1": 1.0,
                         "orth_W": 0.1, "alpha_W": 0.1, "l1_ratio_W": 0.1})
    assert bad["status"] == STATUS_FAIL and bad["loss"] == np.inf   # guided > unguided: optimization.py:185-188


def test_search_dispatches_folds_over_devices_and_builds_history(tmp_path):
    _FakeModel.seen = []
    opt = ComponentOptimizer(_adata(), ["batch", "cond"], max_iter=5, device="cpu", random_state=3)
    opt.model_factory = _FakeModel
    opt.devices = ["cuda:0", "cuda:1", "cuda:2"]     # pretend three GPUs: the stand-in model only records them
    best = opt.search_hyperparams(n_total_components_range=(10, 30), n_splits=3, max_evals=6)
    assert set(best) == {"n_components", "n_covariate_components", "lam", "alpha_W", "orth_W", "l1_ratio_W", "random_state"}
    assert all(isinstance(v, float) for v in best["lam"]) and len(best["n_covariate_components"]) == 2
    ok = [t for t in opt.trials.trials if t["result"]["status"] == STATUS_OK]
    assert len(opt.trials.trials) == 6 and len(ok) >= 1
    assert len(_FakeModel.seen) == 3 * len(ok) and set(_FakeModel.seen) <= set(opt.devices)
    assert {d for _, d in opt.last_scheduler.assignments} <= set(opt.devices)
    hist = opt.get_train_history()
    assert list(hist["score"]) == sorted(hist["score"], reverse=True)
    assert {"n_components", "n_covariate_components_0", "n_covariate_components_1", "n_total_components", "lam_0",
            "lam_1", "tid", "score"} <= set(hist.columns)
    assert opt.get_hyperparameter(0)["score"] == hist["score"].iloc[0]
    path = tmp_path / "trials.pkl"
    opt.save_trials(str(path))
    opt.extend_training(extra_evals=2)
    assert len(opt.trials.trials) == 8
    opt2 = ComponentOptimizer(_adata(), ["batch", "cond"], max_iter=5, device="cpu")
    opt2.model_factory = _FakeModel
    opt2.search_hyperparams(n_total_components_range=(10, 30), n_splits=3, max_evals=1, trials_filename=str(path))
    assert len(opt2.trials.trials) == 7             # resumed from the six saved trials
    assert isinstance(opt.fit_the_best_param(), _FakeModel)


def test_batched_trials_fill_all_devices_and_match_sequential_search():
    """8 GPUs, 3 folds: two trials (6 fold fits) are in flight at a time; suggestions, scores and trial order are
    those of the one-device search."""
    runs = []
    for devices in (["cuda:0"], [f"cuda:{i}" for i in range(8)]):
        _FakeModel.seen = []
        opt = ComponentOptimizer(_adata(), ["batch", "cond"], max_iter=5, device="cpu", random_state=5)
        opt.model_factory = _FakeModel
        opt.devices = devices
        opt.search_hyperparams(n_total_components_range=(10, 30), n_splits=3, max_evals=7)
        runs.append((opt, list(_FakeModel.seen)))
    (one, seen1), (eight, seen8) = runs
    assert [t["misc"]["vals"] for t in one.trials.trials] == [t["misc"]["vals"] for t in eight.trials.trials]
    assert [t["result"]["loss"] for t in one.trials.trials] == [t["result"]["loss"] for t in eight.trials.trials]
    assert one.best_param == eight.best_param
    assert len(seen1) == len(seen8) and len(set(seen8)) > 3     # more GPUs busy than one trial's three folds
    assert set(seen1) == {"cuda:0"}


def test_optimizer_validation():
    ad = _adata()
    with pytest.raises(TypeError):
        ComponentOptimizer("x", ["batch"])
    with pytest.raises(ValueError):
        ComponentOptimizer(ad, ["nope"])
    with pytest.raises(ValueError):
        ComponentOptimizer(ad, ["batch"], loss_type="l2")
    opt = ComponentOptimizer(ad, ["batch"], max_iter=3, device="cpu")
    with pytest.raises(TypeError):
        opt.search_hyperparams(n_total_components_range=[10, 20])
    with pytest.raises(TypeError):
        opt.search_hyperparams(lam_range=(1, 10))
    with pytest.raises(ValueError):
        opt.search_hyperparams(l1_ratio_W_range=(0.0, 1.5))
    with pytest.raises(ValueError):
        opt.search_hyperparams(n_splits=1)
    with pytest.raises(ValueError):
        opt.search_hyperparams(max_evals=0)
