"""alpine_b200: B200-native drop-in for ALPINE's covariate-guided MU-NMF loop.

The public classes mirror the reference package (``alpine/__init__.py:1-3``).
They are imported lazily so that host-side utilities can be used without the
CUDA extension; the fit/transform path itself fails loudly when the extension
or a GPU is missing (there is no CPU fallback).
"""
__all__ = ["ALPINE", "ComponentOptimizer", "AlpineMatrices"]


def __getattr__(name):
    if name in ("ALPINE", "AlpineMatrices"):
        from . import main

        return getattr(main, name)
    if name == "ComponentOptimizer":
        from .optimization import ComponentOptimizer

        return ComponentOptimizer
    raise AttributeError(name)
