"""Epoch / batch index generation, mirroring the reference's ``alpine/utils/sampling.py``.

Same function names and results; the only change is that the joint labels are built with one vectorised argmax per
covariate instead of a Python loop with a ``.item()`` device sync per cell and covariate (sampling.py:36-55).
``generate_epoch_indices("random")`` is the same ``torch.randperm(n, device=device)`` call, so for a given generator
state it yields the reference's permutation; ``"weighted"`` uses the same class-balanced
``torch.utils.data.WeightedRandomSampler`` with replacement (sampling.py:18-33), with sklearn's
``compute_sample_weight("balanced")`` restated as ``n / (n_classes * count(class))``.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch


def create_joint_labels_from_dummy_matrices(Ys: List[torch.Tensor]) -> List[str]:
    """"cov0_label<a>+cov1_label<b>..." per cell; an all-zero (missing) column maps to label 0, as argmax does."""
    total_samples = Ys[0].shape[1]
    codes = [torch.argmax(Y, dim=0).cpu().numpy() for Y in Ys]
    return ["+".join(f"cov{t}_label{int(c[j])}" for t, c in enumerate(codes)) for j in range(total_samples)]


def _balanced_sample_weight(labels: List[str]) -> np.ndarray:
    uniq, inv, counts = np.unique(np.asarray(labels), return_inverse=True, return_counts=True)
    return (len(labels) / (len(uniq) * counts.astype(np.float64)))[inv]


def generate_epoch_indices(joint_labels: List[str], sampling_method: str, device: torch.device, **kwargs) -> torch.Tensor:
    """``kwargs`` (the reference's signature has them too, unused): ``generator`` = the device generator ``randperm``
    draws from, ``cpu_generator`` = the CPU generator of the weighted sampler; ``None`` = the process-wide ones, which
    is what the reference uses."""
    total_samples = len(joint_labels)
    if sampling_method == "weighted":
        from torch.utils.data import WeightedRandomSampler

        sampler = WeightedRandomSampler(weights=_balanced_sample_weight(joint_labels), num_samples=total_samples,
                                        replacement=True, generator=kwargs.get("cpu_generator"))
        return torch.tensor(list(sampler), device=device, dtype=torch.long)
    if sampling_method == "random":
        gen = kwargs.get("generator")
        if gen is not None and gen.device != torch.device(device):
            gen = None
        return torch.randperm(total_samples, device=device, generator=gen)
    raise ValueError(f"Unknown sampling method: {sampling_method}. Only 'weighted', and 'random' are supported.")


def get_batch_indices(epoch_indices: torch.Tensor, batch_num: int, batch_size: int) -> torch.Tensor:
    start = batch_num * batch_size
    if start >= len(epoch_indices):
        return torch.empty(0, device=epoch_indices.device, dtype=torch.long)
    return epoch_indices[start:min(start + batch_size, len(epoch_indices))]


def get_num_batches(total_samples: int, batch_size: int) -> int:
    return (total_samples + batch_size - 1) // batch_size
