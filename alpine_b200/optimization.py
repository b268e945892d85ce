"""Drop-in ``ComponentOptimizer`` (reference ``alpine/optimization.py``), scheduling one fit per GPU.

In scope here: the class and method signatures, the component-split rule, the objective and the
DISPATCH of the independent fold fits over the GPUs of the box (``alpine_b200.scheduler``).  Out of scope and only
bridged: the optimiser itself (hyperopt's TPE) and the clustering used for scoring (scanpy neighbours + Leiden) are
third-party packages that are not installed in this image.  When they are importable they are used exactly as the
reference does (optimization.py:95-130, 271-272); otherwise a seeded random search over the same space and a
k-means clustering of the embedding (scikit-learn) stand in, so that the scheduler can be exercised end to end.
Like the reference, the mean CV score is handed to the optimiser as its "loss" (optimization.py:216, 285) and
``get_train_history`` sorts by it in descending order (optimization.py:473).
"""
from __future__ import annotations

import pickle
import threading
from copy import copy
from typing import Callable, List, Optional, Tuple

import numpy as np
import pandas as pd

from .main import ALPINE
from .scheduler import DeviceScheduler, visible_devices
from .utils.anndata_compat import HAVE_ANNDATA, AnnData

STATUS_OK, STATUS_FAIL = "ok", "fail"

try:  # pragma: no cover - hyperopt is not installed in this image
    import hyperopt as _hyperopt
except Exception:
    _hyperopt = None


# ----------------------------------------------------------------------------------------------------------------
# stand-ins for the third-party pieces
# ----------------------------------------------------------------------------------------------------------------
class Trials:
    """Minimal trial store with the fields the reference reads: ``.trials`` -> [{"tid", "result", "misc"}]."""

    def __init__(self):
        self.trials: List[dict] = []

    def add(self, vals: dict, result: dict) -> None:
        self.trials.append({"tid": len(self.trials), "result": result, "misc": {"vals": vals}})

    def best_vals(self) -> Optional[dict]:
        ok = [t for t in self.trials if t["result"]["status"] == STATUS_OK]
        if not ok:
            return None
        return min(ok, key=lambda t: t["result"]["loss"])["misc"]["vals"]  # the optimiser MINIMISES (as fmin)


def _sample_space(rng: np.random.Generator, ranges: dict, n_cov: int) -> dict:
    """One draw from the reference's search space (optimization.py:95-120): quniform / uniform / qloguniform."""
    lo, hi = ranges["n_total_components"]
    vals = {
        "n_total_components": float(np.round(rng.uniform(lo, hi))),
        "orth_W": float(rng.uniform(*ranges["orth_W"])),
        "alpha_W": float(rng.uniform(*ranges["alpha_W"])),
        "l1_ratio_W": float(rng.uniform(*ranges["l1_ratio_W"])),
    }
    for i in range(n_cov + 1):
        vals[f"split_{i}"] = float(rng.uniform(0, 1))
    llo, lhi = np.log(ranges["lam"][0]), np.log(ranges["lam"][1])
    for i in range(n_cov):
        vals[f"lam_{i}"] = float(np.round(np.exp(rng.uniform(llo, lhi))))
    return vals


def _vals_to_space(vals: dict, n_cov: int) -> dict:
    space = {k: vals[k] for k in ("n_total_components", "orth_W", "alpha_W", "l1_ratio_W")}
    space["splits"] = [vals[f"split_{i}"] for i in range(n_cov + 1)]
    for i in range(n_cov):
        space[f"lam_{i}"] = vals[f"lam_{i}"]
    return space


def default_scorer(val_adata, covariate_keys: List[str], random_state: int) -> float:
    """ARI + homogeneity of a clustering of ``ALPINE_embedding`` against every covariate (optimization.py:271-280).

    Clustering: scanpy neighbours + Leiden (resolution 1) when scanpy is installed, as the reference; otherwise
    k-means with one cluster per joint label (stand-in)."""
    from sklearn.metrics.cluster import adjusted_rand_score, homogeneity_score

    emb = np.asarray(val_adata.obsm["ALPINE_embedding"])
    try:  # pragma: no cover - scanpy is not installed in this image
        import scanpy as sc

        sc.pp.neighbors(val_adata, use_rep="ALPINE_embedding")
        sc.tl.leiden(val_adata, flavor="igraph", resolution=1)
        clusters = np.asarray(val_adata.obs["leiden"])
    except ImportError:
        from sklearn.cluster import KMeans

        n_joint = len(val_adata.obs[covariate_keys].astype(str).drop_duplicates())  # distinct joint labels
        k = int(max(2, min(n_joint, len(emb) - 1)))
        # (fp64 input: scikit-learn's distance kernels otherwise upcast fp32 chunk by chunk, thousands of small copies)
        clusters = KMeans(n_clusters=k, n_init=4, random_state=random_state).fit_predict(emb.astype(np.float64))
    score = 0.0
    for key in covariate_keys:
        keep = ~val_adata.obs[key].isna().to_numpy()
        truth = val_adata.obs[key].to_numpy()[keep].astype(str)
        score += adjusted_rand_score(truth, clusters[keep])
        score += homogeneity_score(truth, clusters[keep])
    return score / len(covariate_keys)


# ----------------------------------------------------------------------------------------------------------------
class ComponentOptimizer:
    def __init__(
        self,
        adata: AnnData,
        covariate_keys: List[str],
        use_als: bool = False,
        loss_type: str = "kl-divergence",
        max_iter: Optional[int] = None,
        batch_size: Optional[int] = None,
        sampling_method: str = "random",
        device: str = "cuda",
        random_state: int = 42,
    ):
        self._validate_init_args(adata, covariate_keys, loss_type, max_iter, batch_size, device, random_state)
        self.adata = adata.copy()
        self.covariate_keys = covariate_keys
        self.use_als = use_als
        self.loss_type = loss_type
        self.max_iter = max_iter
        self.batch_size = batch_size
        self.sampling_method = sampling_method
        self.device = device
        self.random_state = random_state
        self.best_param: dict = {}
        self.max_iter_detect = self.max_iter is None
        if self.max_iter_detect:
            print("Owing to max_iter being None, it will be determine by the average of the first n_splits iterations.")
        # scheduling: one fit per GPU (the reference runs the folds one after the other on one device)
        self.devices: List[str] = visible_devices(device)
        self._fold_cache: dict = {}
        self.device_busy_s: dict = {}  # device -> seconds spent inside fold jobs, summed over the whole search
        self._fold_lock = threading.Lock()
        self.scorer: Callable = default_scorer
        self.model_factory: Callable[..., ALPINE] = ALPINE

    # ------------------------------------------------------------------------------------------------ search
    def search_hyperparams(
        self,
        n_total_components_range: Tuple[int, int] = (10, 100),
        lam_range: Tuple[float, float] = (1.0, 1e4),
        orth_W_range: Tuple[float, float] = (0.0, 1.0),
        alpha_W_range: Tuple[float, float] = (0.0, 100.0),
        l1_ratio_W_range: Tuple[float, float] = (0.0, 1.0),
        min_covariate_components: Optional[List[int]] = None,
        n_splits: int = 3,
        max_evals: int = 100,
        trials_filename: Optional[str] = None,
    ):
        self._validate_search_args(n_total_components_range, lam_range, orth_W_range, alpha_W_range,
                                   l1_ratio_W_range, n_splits, max_evals)
        self.iter_records: List[int] = []
        self.n_splits = n_splits
        if trials_filename is not None:
            self.load_trials(trials_filename)
        else:
            self.trials = _hyperopt.Trials() if _hyperopt is not None else Trials()
        if min_covariate_components is None:
            self.min_covariate_components = [self.adata.obs[key].nunique() for key in self.covariate_keys]
        else:
            if isinstance(min_covariate_components, list) and len(min_covariate_components) != len(self.covariate_keys):
                raise ValueError("min_covariate_components should have the same length as the number of covariates.")
            if any(comp < 2 for comp in min_covariate_components):
                raise ValueError("min_covariate_components should be greater than or equal to 2.")
            self.min_covariate_components = min_covariate_components
        self._ranges = {"n_total_components": n_total_components_range, "lam": lam_range, "orth_W": orth_W_range,
                        "alpha_W": alpha_W_range, "l1_ratio_W": l1_ratio_W_range}
        best = self._run_search(max_evals)
        return self._store_best(best)

    def extend_training(self, extra_evals: int = 50):
        if not hasattr(self, "trials"):
            raise RuntimeError("Please run bayesian_search() before extending training.")
        best = self._run_search(extra_evals)
        self._store_best(best)
        return copy(self.best_param)

    def _run_search(self, n_new: int) -> dict:
        n_cov = len(self.covariate_keys)
        if _hyperopt is not None:  # pragma: no cover - exactly the reference's call (optimization.py:95-130)
            hp, r = _hyperopt.hp, self._ranges
            self.space = {
                "n_total_components": hp.quniform("n_total_components", r["n_total_components"][0], r["n_total_components"][1], 1),
                "orth_W": hp.uniform("orth_W", *r["orth_W"]),
                "alpha_W": hp.uniform("alpha_W", *r["alpha_W"]),
                "l1_ratio_W": hp.uniform("l1_ratio_W", *r["l1_ratio_W"]),
                "splits": [hp.uniform(f"split_{i}", 0, 1) for i in range(n_cov + 1)],
            }
            for i in range(n_cov):
                self.space[f"lam_{i}"] = hp.qloguniform(f"lam_{i}", np.log(r["lam"][0]), np.log(r["lam"][1]), 1)
            best = _hyperopt.fmin(self.objective, self.space, algo=_hyperopt.tpe.suggest,
                                  max_evals=n_new + len(self.trials.trials), trials=self.trials,
                                  rstate=np.random.default_rng(self.random_state))
        else:
            # Random search: the suggestions do not depend on earlier results, so one trial per GPU is evaluated at
            # a time and all their fold fits go through ONE device queue (8 GPUs: 8 trials x 3 folds = 24 jobs, the
            # queue keeps every GPU busy until the batch drains).  Same suggestions, same scores and the same trial
            # order as evaluating them one by one.
            rng = np.random.default_rng(self.random_state + len(self.trials.trials))
            batch = max(1, len(self.devices))
            todo = n_new
            while todo > 0:
                vals = [_sample_space(rng, self._ranges, n_cov) for _ in range(min(batch, todo))]
                for v, result in zip(vals, self._evaluate([_vals_to_space(v, n_cov) for v in vals])):
                    self.trials.add(v, result)
                todo -= len(vals)
            best = self.trials.best_vals()
        if best is None:
            raise RuntimeError("Hyperparameter optimization did not return any result.")
        return best

    def _store_best(self, best: dict) -> dict:
        n_cov = len(self.covariate_keys)
        n_components, n_covariate_components = self._distribute_components(
            {"n_total_components": best["n_total_components"], "splits": [best[f"split_{i}"] for i in range(n_cov + 1)]})
        self.best_param["n_components"] = n_components
        self.best_param["n_covariate_components"] = n_covariate_components
        self.best_param["lam"] = [float(best[f"lam_{i}"]) for i in range(n_cov)]
        self.best_param["alpha_W"] = best["alpha_W"]
        self.best_param["orth_W"] = best["orth_W"]
        self.best_param["l1_ratio_W"] = best["l1_ratio_W"]
        self.best_param["random_state"] = self.random_state
        return self.best_param

    def _distribute_components(self, space):
        """Half of the total goes to the unguided block, the rest is split by the normalised ratios and floored at
        ``min_covariate_components`` (optimization.py:153-176)."""
        total = int(space["n_total_components"])
        ratios = np.array([float(s) for s in space["splits"]])
        ratios = ratios / np.sum(ratios)
        rest = total - int(total / 2)
        guided = [int(round(rest * r)) for r in ratios[:-1]]
        guided = [max(self.min_covariate_components[i], g) for i, g in enumerate(guided)]
        return total - sum(guided), guided

    def objective(self, space):
        return self._evaluate([space])[0]

    def _trial_args(self, space) -> Optional[dict]:
        """Model arguments of one suggestion, or None when the component split is infeasible (optimization.py:185-188)."""
        n_cov = len(self.covariate_keys)
        lam = [space[f"lam_{i}"] for i in range(n_cov)]
        n_components, n_covariate_components = self._distribute_components(space)
        if not (sum(n_covariate_components) <= n_components and all(n >= 2 for n in n_covariate_components)):
            return None
        return {"n_components": n_components, "n_covariate_components": n_covariate_components, "lam": lam,
                "orth_W": space["orth_W"], "alpha_W": space["alpha_W"], "l1_ratio_W": space["l1_ratio_W"]}

    def _evaluate(self, spaces: List[dict]) -> List[dict]:
        """Objective of one or several suggestions (optimization.py:178-218); all their fold fits share one device
        queue.  With ``max_iter=None`` every trial of a batch detects its own iteration count; the running mean is
        taken over after the batch instead of after the first trial."""
        all_args = [self._trial_args(sp) for sp in spaces]
        folds = self._folds()
        jobs = [(a, tr, va) for a in all_args if a is not None for tr, va in folds]
        self.last_scheduler = DeviceScheduler(self.devices)
        out = self.last_scheduler.map(self._fit_fold, jobs) if jobs else []
        self._account_busy()
        results, pos = [], 0
        for a in all_args:
            if a is None:
                results.append({"loss": np.inf, "status": STATUS_FAIL})  # optimization.py:217-218
                continue
            mine = out[pos:pos + len(folds)]
            pos += len(folds)
            if self.max_iter_detect:
                self.iter_records.extend(m for _, m in mine)
            score = float(np.mean([s for s, _ in mine]))
            params = dict(a, lam=list(a["lam"]),
                          max_iter=self.iter_records[-1] if self.max_iter_detect else self.max_iter, score=score)
            results.append({"loss": score, "status": STATUS_OK, "params": params})
        if self.max_iter_detect and len(self.iter_records) >= self.n_splits:
            self.max_iter = int(sum(self.iter_records) / len(self.iter_records))
        return results

    # ------------------------------------------------------------------------------------------------ scoring
    def _folds(self):
        if len(self.covariate_keys) == 1:
            joint = self.adata.obs[self.covariate_keys[0]].astype(str)
        else:
            joint = self.adata.obs[self.covariate_keys[0]].astype(str)
            for key in self.covariate_keys[1:]:
                joint = joint + "_" + self.adata.obs[key].astype(str)
        from sklearn.model_selection import StratifiedKFold

        skf = StratifiedKFold(n_splits=self.n_splits, shuffle=True, random_state=self.random_state)
        return list(skf.split(np.zeros(len(joint)), joint))

    def _fold_adata(self, idx) -> AnnData:
        """``self.adata[idx].copy()`` (optimization.py:242-243).  The folds are the same for every trial, so the
        row subsets of X and obs are materialised once per fold and every job gets its own AnnData around them
        (fresh obsm / varm / layers, its own obs frame): fits never write to X, and concurrent trials must not share
        the slots they do write."""
        key = np.ascontiguousarray(idx, dtype=np.int64).tobytes()  # the index content itself: no collisions
        with self._fold_lock:
            hit = self._fold_cache.get(key)
            if hit is None:
                sub = self.adata[idx]
                sub = sub.copy() if HAVE_ANNDATA else sub  # the stand-in's subsetting already copies
                hit = self._fold_cache[key] = sub
        if HAVE_ANNDATA:  # pragma: no cover - real anndata: keep the reference's semantics literally
            return hit.copy()
        return AnnData(hit.X, obs=hit.obs.copy(), var=hit.var)

    def _fit_fold(self, job, device: str):
        """One fold: fit on the training cells, transform the validation cells, score the embedding."""
        args, train_idx, val_idx = job
        train_adata = self._fold_adata(train_idx)
        val_adata = self._fold_adata(val_idx)
        model = self.model_factory(
            n_covariate_components=args["n_covariate_components"], n_components=args["n_components"],
            lam=[float(v) for v in args["lam"]], orth_W=float(args["orth_W"]), alpha_W=float(args["alpha_W"]),
            l1_ratio_W=float(args["l1_ratio_W"]), use_als=self.use_als, random_state=self.random_state,
            loss_type=self.loss_type, device=device)
        # worker threads run concurrently: the model keeps its random streams to itself instead of publishing them to
        # the process-wide generators (same numbers as the reference's sequential fit -> transform on one device)
        model._rng_private = True
        model.fit(adata=train_adata, covariate_keys=self.covariate_keys, max_iter=self.max_iter,
                  batch_size=self.batch_size, sampling_method=self.sampling_method, verbose=False)
        # (the reference stores the embeddings into train_adata again here, optimization.py:267; fit has just done
        # that and train_adata is dropped, so the call is skipped)
        model.transform(val_adata)
        return self.scorer(val_adata, self.covariate_keys, self.random_state), model.max_iter

    def calc_score(self, args):
        """Mean CV score of one hyper-parameter setting (optimization.py:220-287); the folds run concurrently, one
        per GPU."""
        jobs = [(args, tr, va) for tr, va in self._folds()]
        self.last_scheduler = DeviceScheduler(self.devices)
        out = self.last_scheduler.map(self._fit_fold, jobs)
        self._account_busy()
        if self.max_iter_detect:
            self.iter_records.extend(m for _, m in out)
        return float(np.mean([s for s, _ in out]))

    def _account_busy(self) -> None:
        for d, t in self.last_scheduler.busy_s.items():
            self.device_busy_s[d] = self.device_busy_s.get(d, 0.0) + t

    # ------------------------------------------------------------------------------------------- persistence
    def save_trials(self, filename: str):
        with open(filename, "wb") as f:
            pickle.dump(self.trials, f)
        print(f"Trials saved to {filename}")

    def load_trials(self, filename: str):
        with open(filename, "rb") as f:
            self.trials = pickle.load(f)
        print(f"Trials loaded from {filename}")

    def get_hyperparameter(self, idx):
        tid = self.get_train_history().iloc[idx]["tid"]
        for trial in self.trials.trials:
            if trial["tid"] == tid:
                return trial["result"]["params"]

    def get_train_history(self):
        """One row per successful trial, list-valued parameters expanded, sorted by score descending."""
        rows = []
        for trial in self.trials.trials:
            if "result" in trial and trial["result"]["status"] == STATUS_OK:
                info = dict(trial["result"]["params"])
                info["score"] = trial["result"]["loss"]
                info["tid"] = trial["tid"]
                rows.append(info)
        df = pd.DataFrame(rows)
        n_cov = len(df["n_covariate_components"].iloc[0])
        cov_cols = [f"n_covariate_components_{i}" for i in range(n_cov)]
        cov_df = pd.DataFrame(df["n_covariate_components"].tolist(), columns=cov_cols)
        lam_df = pd.DataFrame(df["lam"].tolist(), columns=[f"lam_{i}" for i in range(len(df["lam"].iloc[0]))])
        df = pd.concat([df.drop(columns=["n_covariate_components", "lam"]), cov_df, lam_df], axis=1)
        df["n_total_components"] = df["n_components"] + df[cov_cols].sum(axis=1)
        head = ["n_components"] + cov_cols + ["n_total_components"]
        df = df[head + [c for c in df.columns if c not in head]]
        return df.sort_values(by="score", ascending=False).reset_index(drop=True)

    def fit_the_best_param(self):
        if not getattr(self, "best_param", None):
            raise RuntimeError("Please run bayesian_search() to find the best parameters first.")
        model = self.model_factory(**self.best_param, use_als=self.use_als, loss_type=self.loss_type,
                                   device=self.devices[0] if self.device == "cuda" else self.device)
        model.fit(adata=self.adata, covariate_keys=self.covariate_keys, max_iter=self.max_iter,
                  batch_size=self.batch_size, verbose=False)
        return model

    # -------------------------------------------------------------------------------------------- validation
    def _validate_init_args(self, adata, covariate_keys, loss_type, max_iter, batch_size, device, random_state) -> None:
        """optimization.py:512-550."""
        if not isinstance(adata, AnnData):
            raise TypeError("adata must be an instance of AnnData")
        if not isinstance(covariate_keys, list):
            raise TypeError("covariate_keys must be a list")
        if not all(isinstance(key, str) for key in covariate_keys):
            raise TypeError("All covariate_keys must be strings")
        if not all(key in adata.obs.columns for key in covariate_keys):
            raise ValueError("All covariate_keys must be present in adata.obs")
        if loss_type not in ["kl-divergence", "frobenius"]:
            raise ValueError("loss_type must be either 'kl-divergence' or 'frobenius'")
        for value, name in ((max_iter, "max_iter"), (batch_size, "batch_size")):
            if value is not None and (not isinstance(value, int) or value < 0):
                raise ValueError(f"{name} must be a non-negative integer")
        if not isinstance(random_state, int):
            raise TypeError("random_state must be an integer")

    def _validate_search_args(self, n_total_components_range, lam_range, orth_W_range, alpha_W_range,
                              l1_ratio_W_range, n_splits, max_evals) -> None:
        """optimization.py:552-604."""
        if not isinstance(n_total_components_range, tuple) or len(n_total_components_range) != 2:
            raise TypeError("n_total_components_range must be a tuple of two integers")
        if n_total_components_range[0] >= n_total_components_range[1]:
            raise ValueError("n_total_components_range must be a tuple with the first element less than the second")
        if n_total_components_range[0] < 2:
            raise ValueError("n_total_components_range must be a tuple with the first element greater than or equal to 2")
        for arg, name in ((lam_range, "lam_range"), (orth_W_range, "orth_W_range"), (alpha_W_range, "alpha_W_range"),
                          (l1_ratio_W_range, "l1_ratio_W_range")):
            if not isinstance(arg, tuple) or len(arg) != 2:
                raise TypeError(f"{name} must be a tuple of two floats")
            if not all(isinstance(x, float) for x in arg):
                raise TypeError(f"All elements of {name} must be floats")
            if arg[0] >= arg[1]:
                raise ValueError(f"{name} must be a tuple with the first element less than the second")
        if l1_ratio_W_range[1] > 1.0:
            raise ValueError("l1_ratio_W_range's second element must be less than or equal to 1.0")
        if not isinstance(n_splits, int):
            raise TypeError("n_splits must be an integer")
        if n_splits < 2:
            raise ValueError("n_splits must be greater than or equal to 2")
        if not isinstance(max_evals, int) or max_evals <= 0:
            raise ValueError("max_evals must be a positive integer")
