"""Elbow detection for the ``max_iter=None`` warm-up (reference main.py:755-770).

The reference calls ``kneed.KneeLocator(x, y, curve="convex", direction="decreasing",
interp_method="polynomial", polynomial_degree=2)`` and reads ``.elbow``.  ``kneed`` is a third-party dependency
that is not vendored in the reference tree and is not installed in this image, so the Kneedle procedure
(Satopaa et al., "Finding a 'Kneedle' in a Haystack", 2011) is restated here for exactly that configuration
(sensitivity S = 1, offline mode).  The package cannot be run here; the procedure is pinned on the published
example vectors instead (the paper's Figure 2 curve and the four ``DataGenerator`` shape vectors of the package's
README / test-suite with their documented knees), see tests/test_kneedle_published.py.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def _unit(v: np.ndarray) -> np.ndarray:
    span = v.max() - v.min()
    return (v - v.min()) / span if span > 0 else np.zeros_like(v)


def _plateau_extrema(d: np.ndarray, greater: bool) -> np.ndarray:
    """Indices i (interior) with d[i] >= both neighbours (or <=), ties included, as argrelextrema(order=1)."""
    if d.size < 3:
        return np.empty(0, dtype=np.int64)
    mid, left, right = d[1:-1], d[:-2], d[2:]
    mask = (mid >= left) & (mid >= right) if greater else (mid <= left) & (mid <= right)
    return np.nonzero(mask)[0] + 1


def find_elbow(x, y, sensitivity: float = 1.0, degree: int = 2, curve: str = "convex",
               direction: str = "decreasing", interp: str = "polynomial") -> Optional[float]:
    """x-value of the knee / elbow, or ``None`` if the difference curve never drops below its threshold.

    The defaults are the reference's configuration (convex, decreasing, degree-2 polynomial smoothing).  The other
    curve / direction combinations and ``interp="none"`` (the package's ``interp1d`` mode evaluated on its own sample
    points, i.e. no smoothing) exist so that the procedure can be pinned on the published example vectors of the
    Kneedle paper and of the ``kneed`` package (tests/test_kneedle_published.py)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if x.size < 3:
        return None
    if interp == "polynomial":
        smooth = np.polyval(np.polyfit(x, y, degree), x)      # 1. least-squares smoothing
    elif interp == "none":
        smooth = y
    else:
        raise ValueError("interp must be 'polynomial' or 'none'")
    xn, yn = _unit(x), _unit(smooth)                          # 2. unit square
    # 3. every case becomes "increasing + concave"; two of them by reversing the curve, which maps index i to n-1-i
    flipped = False
    if direction == "decreasing" and curve == "convex":
        yn = yn.max() - yn
    elif direction == "decreasing" and curve == "concave":
        yn, flipped = yn[::-1], True
    elif direction == "increasing" and curve == "convex":
        yn, flipped = (yn.max() - yn)[::-1], True
    elif not (direction == "increasing" and curve == "concave"):
        raise ValueError("curve must be 'convex' or 'concave', direction 'increasing' or 'decreasing'")
    diff = yn - xn                                            #    difference curve
    peaks = _plateau_extrema(diff, greater=True)              # 4. local maxima / minima of the difference curve
    dips = _plateau_extrema(diff, greater=False)
    if peaks.size == 0:
        return None
    thresholds = diff[peaks] - sensitivity * abs(float(np.mean(np.diff(xn))))  # 5. one threshold per maximum
    is_peak = np.zeros(x.size, dtype=bool)
    is_peak[peaks] = True
    is_dip = np.zeros(x.size, dtype=bool)
    is_dip[dips] = True
    threshold, candidate, seen = 0.0, 0, 0
    for i in range(int(peaks[0]), x.size - 1):                # 6. first drop below the active threshold
        if xn[i] == 1.0:
            break
        if is_peak[i]:
            threshold, candidate = float(thresholds[seen]), i
            seen += 1
        if is_dip[i]:
            threshold = 0.0
        if diff[i + 1] < threshold:
            return float(x[x.size - 1 - candidate] if flipped else x[candidate])
    return None
