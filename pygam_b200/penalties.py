"""Penalty and constraint matrices of the m x m system (host side, numpy).

These are the generators of ``/root/reference/pygam/penalties.py`` restated as dense numpy
arrays: ``derivative`` (9-49), ``periodic`` (52-53), ``l2`` (56-73), ``monotonic_inc/dec``
(76-150), ``convex/concave`` (153-258), ``none`` (261-276).  They are O(m^2) objects that the
reference also builds on the host; the n-sized work is on the GPU.
A penalty is any callable ``(n, coef) -> (n, n)`` array, as in the reference
(``terms.py:344-347``).
"""

import numpy as np


def _differences(n, order):
    """columns are the `order`-th forward differences of the n coordinates: shape (n, n-order)"""
    D = np.eye(n)
    for _ in range(order):
        D = D[:, 1:] - D[:, :-1]
    return D


def derivative(n, coef=None, derivative=2, periodic=False):
    if n == 1:
        return np.zeros((1, 1))
    pad = 2 * derivative if periodic else 0
    D = _differences(n + pad, derivative)
    if periodic:
        # wrap the stencil around the ends, as penalties.py:33-48 does
        wrapped = D[:, :derivative].copy()
        D[:, -2 * derivative:-derivative] += wrapped * (-1) ** derivative
        half = (n + pad) // 2
        D[-half:] = D[:half][::-1, ::-1]
        D = D[derivative:-derivative, derivative:-derivative]
    return D @ D.T


def periodic(n, coef=None, derivative_order=2):
    return derivative(n, coef, derivative=derivative_order, periodic=True)


def l2(n, coef=None):
    return np.eye(n)


def none(n, coef=None):
    return np.zeros((n, n))


def _violation_penalty(n, coef, order, negative_is_violation):
    coef = np.ravel(coef)
    if len(coef) != n:
        raise ValueError(f"dimension mismatch: expected n equals len(coef), but found n = {n}, "
                         f"coef.shape = {coef.shape}.")
    if n == 1:
        return np.zeros((1, 1))
    d = np.diff(coef, n=order)
    viol = (d < 0) if negative_is_violation else (d > 0)
    D = _differences(n, order) * viol.astype(np.float64)[None, :]
    return D @ D.T


def monotonic_inc(n, coef):
    return _violation_penalty(n, coef, 1, True)


def monotonic_dec(n, coef):
    return _violation_penalty(n, coef, 1, False)


def convex(n, coef):
    return _violation_penalty(n, coef, 2, True)


def concave(n, coef):
    return _violation_penalty(n, coef, 2, False)


PENALTIES = {"auto": "auto", "derivative": derivative, "l2": l2, "none": none, "periodic": periodic}
CONSTRAINTS = {"convex": convex, "concave": concave, "monotonic_inc": monotonic_inc,
               "monotonic_dec": monotonic_dec, "none": none}
