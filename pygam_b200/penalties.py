"""Penalty and constraint matrices of the m x m system (host side, numpy).

These are the generators of ``/root/reference/pygam/penalties.py`` restated with dense numpy
arithmetic: ``derivative`` (9-49), ``periodic`` (52-53), ``l2`` (56-73), ``monotonic_inc/dec``
(76-150), ``convex/concave`` (153-258), ``none`` (261-276), ``wrap_penalty`` (279-309), ``sparse_diff``
(312-348).  They are O(m^2) objects that the reference also builds on the host;
the n-sized work is on the GPU.  A penalty is any callable ``(n, coef) -> (n, n)`` matrix, dense or scipy
sparse, as in the reference (``terms.py:344-347``); the public generators return scipy sparse matrices like
the reference's (its tests call ``.toarray()`` on them), the terms densify whatever they get.
"""

import warnings

import numpy as np
import scipy.sparse


def _differences(n, order):
    """columns are the `order`-th forward differences of the n coordinates: sparse (n, n - order), integer entries"""
    D = scipy.sparse.identity(n, format="csc")
    for _ in range(order):
        if D.shape[1] == 0:
            break
        D = (D[:, 1:] - D[:, :-1]).tocsc()
    return D


def derivative(n, coef=None, derivative=2, periodic=False):
    if n == 1:
        return scipy.sparse.csc_matrix((1, 1))
    if not periodic:
        D = _differences(n, derivative)
        return (D @ D.T).tocsc()
    pad = 2 * derivative
    D = _differences(n + pad, derivative).toarray()
    # wrap the stencil around the ends, as penalties.py:33-48 does (a spline term's own penalty: n is small)
    wrapped = D[:, :derivative].copy()
    D[:, -2 * derivative:-derivative] += wrapped * (-1) ** derivative
    half = (n + pad) // 2
    D[-half:] = D[:half][::-1, ::-1]
    D = D[derivative:-derivative, derivative:-derivative]
    return scipy.sparse.csc_matrix(D @ D.T)


def periodic(n, coef=None, derivative_order=2):
    return derivative(n, coef, derivative=derivative_order, periodic=True)


def l2(n, coef=None):
    return scipy.sparse.identity(n, format="csc")


def none(n, coef=None):
    return scipy.sparse.csc_matrix((n, n))


def _violation_penalty(n, coef, order, negative_is_violation):
    coef = np.ravel(coef)
    if len(coef) != n:
        raise ValueError(f"dimension mismatch: expected n equals len(coef), but found n = {n}, "
                         f"coef.shape = {coef.shape}.")
    if n == 1:
        return scipy.sparse.csc_matrix((1, 1))
    d = np.diff(coef, n=order)
    viol = (d < 0) if negative_is_violation else (d > 0)
    D = _differences(n, order) @ scipy.sparse.diags(viol.astype(np.float64), format="csc")
    return (D @ D.T).tocsc()


def monotonic_inc(n, coef):
    return _violation_penalty(n, coef, 1, True)


def monotonic_dec(n, coef):
    return _violation_penalty(n, coef, 1, False)


def convex(n, coef):
    return _violation_penalty(n, coef, 2, True)


def concave(n, coef):
    return _violation_penalty(n, coef, 2, False)


# ---- the same matrices as dense numpy arrays: what the terms hand to the engine (m x m, m <= a few thousand).  Building
# ---- them through scipy.sparse costs ~0.3 ms per call, which a small fit or a grid-search candidate would notice.
def _differences_dense(n, order):
    D = np.eye(n)
    for _ in range(order):
        D = D[:, 1:] - D[:, :-1]
    return D


def _derivative_dense(n, coef=None, derivative=2, periodic=False):
    if n == 1:
        return np.zeros((1, 1))
    if periodic:
        return globals()["derivative"](n, coef, derivative=derivative, periodic=True).toarray()
    D = _differences_dense(n, derivative)
    return D @ D.T


def _violation_dense(n, coef, order, negative_is_violation):
    coef = np.ravel(coef)
    if len(coef) != n:
        raise ValueError(f"dimension mismatch: expected n equals len(coef), but found n = {n}, "
                         f"coef.shape = {coef.shape}.")
    if n == 1:
        return np.zeros((1, 1))
    d = np.diff(coef, n=order)
    viol = (d < 0) if negative_is_violation else (d > 0)
    D = _differences_dense(n, order) * viol.astype(np.float64)[None, :]
    return D @ D.T


DENSE = {derivative: _derivative_dense,
         periodic: lambda n, coef=None: _derivative_dense(n, coef, 2, True),
         l2: lambda n, coef=None: np.eye(n),
         none: lambda n, coef=None: np.zeros((n, n)),
         monotonic_inc: lambda n, coef: _violation_dense(n, coef, 1, True),
         monotonic_dec: lambda n, coef: _violation_dense(n, coef, 1, False),
         convex: lambda n, coef: _violation_dense(n, coef, 2, True),
         concave: lambda n, coef: _violation_dense(n, coef, 2, False)}


def dense(fn, n, coef=None):
    """the (n, n) matrix of a penalty / constraint generator as a dense array: the built-in ones directly, any other
    callable through its own return value (dense or scipy sparse, terms.py:344-347)"""
    if fn in DENSE:
        return DENSE[fn](n, coef)
    out = fn(n, coef)
    return np.asarray(out.toarray() if hasattr(out, "toarray") else out, dtype=np.float64)


def wrap_penalty(p, fit_linear, linear_penalty=0.0):
    """penalties.py:279-309: a penalty generator whose first coefficient (the linear term of the feature) gets
    ``linear_penalty`` and whose other n - 1 coefficients get ``p``"""

    def wrapped_p(n, *args):
        if not fit_linear:
            return p(n, *args)
        if n == 1:
            return scipy.sparse.block_diag([linear_penalty], format="csc")
        return scipy.sparse.block_diag([linear_penalty, p(n - 1, *args)], format="csc")

    return wrapped_p


def sparse_diff(array, n=1, axis=-1):
    """penalties.py:312-348: np.diff for a scipy sparse matrix (order n along axis 0, 1 or -1)"""
    if n < 0 or int(n) != n:
        raise ValueError(f"Expected order is non-negative integer, but found: {n}")
    if not scipy.sparse.issparse(array):
        warnings.warn(f"Array is not sparse. Consider using numpy.diff")
    if n == 0:
        return array
    nd = array.ndim
    if axis < 0:
        axis += nd
    A = scipy.sparse.csc_matrix(array)
    for _ in range(int(n)):
        A = (A[:, 1:] - A[:, :-1]) if axis == 1 else (A[1:, :] - A[:-1, :])
    return A


PENALTIES = {"auto": "auto", "derivative": derivative, "l2": l2, "none": none, "periodic": periodic}
CONSTRAINTS = {"convex": convex, "concave": concave, "monotonic_inc": monotonic_inc,
               "monotonic_dec": monotonic_dec, "none": none}
