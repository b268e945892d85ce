"""Pin the Kneedle restatement (alpine_b200/utils/kneedle.py, the reference's main.py:755-770 call into the absent
``kneed`` package) on PUBLISHED example vectors instead of on its own twin in oracle/:

* the curve of Figure 2 of Satopaa et al., "Finding a 'Kneedle' in a Haystack" (2011): y = -1/(x + 0.1) + 5 on ten
  points of [0, 1]; the paper and the package's README (``DataGenerator.figure2()``) give the knee at x = 0.22;
* the four shape vectors of the package's ``DataGenerator`` (``convex_increasing`` ... ``concave_decreasing``, x =
  0..9) with the knees its README / test-suite document: 7, 2, 2, 7.

These run the package's default ``interp1d`` mode, which on its own sample points is the identity (``interp="none"``
here); the reference's configuration only swaps the smoother for a degree-2 polynomial fit, covered by the oracle test.
"""
import math

import numpy as np
import pytest

from alpine_b200.utils.kneedle import find_elbow

Y_CONVEX_INC = np.array([1, 2, 3, 4, 5, 10, 15, 20, 40, 100], dtype=float)


def test_figure2_of_the_kneedle_paper():
    x = np.linspace(0.0, 1.0, 10)
    y = np.true_divide(-1, x + 0.1) + 5
    knee = find_elbow(x, y, curve="concave", direction="increasing", interp="none")
    assert math.isclose(knee, 0.22, rel_tol=0.05)
    assert knee == pytest.approx(2.0 / 9.0)


@pytest.mark.parametrize("curve,direction,y,knee", [
    ("convex", "increasing", Y_CONVEX_INC, 7),
    ("convex", "decreasing", Y_CONVEX_INC[::-1], 2),
    ("concave", "decreasing", 100 - Y_CONVEX_INC, 7),
    ("concave", "increasing", 100 - Y_CONVEX_INC[::-1], 2),
])
def test_data_generator_shape_vectors(curve, direction, y, knee):
    assert find_elbow(np.arange(10), y, curve=curve, direction=direction, interp="none") == knee


def test_reference_configuration_on_a_loss_like_curve():
    """Convex, decreasing, degree-2 polynomial smoothing (main.py:758-765): an exponential decay onto a plateau has
    its elbow where the smoothed difference curve peaks; the answer must be an x of the input grid."""
    x = np.arange(200)
    y = np.log10(1e6 * np.exp(-x / 12.0) + 5e4)
    elbow = find_elbow(x, y)
    assert elbow is not None and elbow == int(elbow) and 20 <= elbow <= 120
