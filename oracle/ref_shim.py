"""Import shim that makes the unmodified reference importable in this container.

TEST INFRASTRUCTURE ONLY (see oracle/alpine_oracle.py header).  The reference
(``/root/reference/alpine/main.py:6-10``, ``optimization.py:10``) imports
``anndata``, ``scanpy``, ``kneed`` and ``hyperopt``, none of which is installed
here and none of which the MU loop uses.  Registering empty stand-in modules
lets ``alpine.main`` import; ``ALPINE._initialize_matrices``, ``_fit``,
``_scale_matrices`` and ``_compute_loss`` then run unmodified on CPU.

``/root/reference`` exists only in the build container, never on the GPU box,
so this module is used exclusively by ``oracle/gen_golden.py`` (fixture
generation) and by CPU tests that skip when the directory is absent.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ALPINE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "alpine", "main.py"))


def _stub(name: str, **attrs) -> None:
    if name in sys.modules:
        return
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod


def import_reference():
    """Return the reference's ``alpine.main`` module (unmodified source)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    _stub("anndata", AnnData=type("AnnData", (), {}))
    _stub("scanpy")
    _stub("kneed", KneeLocator=object)
    _stub("hyperopt", fmin=None, tpe=None, hp=None, Trials=object, STATUS_OK="ok", STATUS_FAIL="fail")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    return importlib.import_module("alpine.main")


def make_reference_model(ref_main, n_cells: int, covariate_keys, max_iter: int = 1, **kwargs):
    """Construct the reference ``ALPINE`` and set what ``fit`` would set (main.py:95-99, 112, 131)."""
    m = ref_main.ALPINE(device="cpu", **kwargs)
    m.covariate_keys = list(covariate_keys)
    m.sampling_method = "random"
    m.verbose = False
    m.batch_size = n_cells
    m.max_iter = max_iter
    return m
