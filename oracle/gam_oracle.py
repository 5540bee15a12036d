"""CPU oracle for the pyGAM PIRLS hot path -- TEST INFRASTRUCTURE ONLY.

This module is a plain numpy/scipy restatement of the reference algorithm
(dswah/pyGAM 0.12.0, ``/root/reference/pygam``) for the path named in BASELINE.json:
model-matrix construction, one PIRLS iteration, the statistics pass and the lam grid
search.  It is the *checker* for the CUDA engine in ``pygam_b200``; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product never does, and has no CPU fallback.

Parity pin: every function here is checked against outputs of the unmodified reference
(``tests/golden/*.npz``, produced by ``tests/golden/make_golden.py`` in the authoring
container) and against the reference's published known answers in
``tests/test_oracle_golden.py``.

It deliberately keeps the reference's *numerical method* (dense QR of W.B followed by an
SVD of [R; E]) so that it doubles as the CPU baseline; the product replaces that method
with a fused Gram accumulation and a null-space-aware m x m solve.

Each function cites the reference lines it restates.  Models are described by plain
dicts (see ``tests/golden/cases.py``), not by the reference's classes.
"""

from __future__ import annotations

import copy

import numpy as np
import scipy.linalg
import scipy.sparse as sps
import scipy.stats

EPS = np.finfo(np.float64).eps  # pygam/utils.py:16

CONSTRAINT_LAM = 1e9  # pygam/pygam.py:187
CONSTRAINT_L2 = 1e-3  # pygam/pygam.py:188
CONSTRAINT_L2_MAX = 1e-1  # pygam/pygam.py:189


class NotPositiveDefiniteError(ValueError):  # pygam/utils.py:22-23
    pass


class OptimizationError(ValueError):  # pygam/utils.py:26-27
    pass


# ======================================================================================
# term descriptions
# ======================================================================================
def _as_list(v):
    if isinstance(v, (list, tuple, np.ndarray)):
        return list(v)
    return [v]


def _norm_spline(d, default_n_splines=20):
    """defaults of ``SplineTerm.__init__`` (terms.py:810-823) + argument normalisation
    (terms.py:148-210, 847-888)"""
    t = dict(kind="s", feature=int(d["feature"]), n_splines=int(d.get("n_splines", default_n_splines)),
             spline_order=int(d.get("spline_order", 3)), basis=d.get("basis", "ps"),
             by=d.get("by"), dtype=d.get("dtype", "numerical"),
             edge_knots=None if d.get("edge_knots") is None else np.asarray(d["edge_knots"], float))
    pens = _as_list(d.get("penalties", "auto"))
    lam = _as_list(d.get("lam", 0.6))
    if len(lam) == 1:
        lam = lam * len(pens)
    if len(lam) != len(pens):
        raise ValueError("expected 1 lam per penalty")
    t["penalties"], t["lam"] = pens, [float(v) for v in lam]
    t["constraints"] = _as_list(d.get("constraints", None))
    if t["basis"] not in ("ps", "cp"):
        raise ValueError("basis must be one of ['ps', 'cp']")
    if not t["n_splines"] > t["spline_order"]:
        raise ValueError("n_splines must be > spline_order")
    return t


def _norm_linear(d):
    """terms.py:662-678"""
    pens = _as_list(d.get("penalties", "auto"))
    lam = _as_list(d.get("lam", 0.6))
    if len(lam) == 1:
        lam = lam * len(pens)
    return dict(kind="l", feature=int(d["feature"]), penalties=pens, lam=[float(v) for v in lam], constraints=[None],
                dtype="numerical", edge_knots=None)


def normalize_terms(term_specs, fit_intercept=True):
    """Canonical term list.  Mirrors the constructors in terms.py (s: 810-845, l: 662-678,
    f: 1010-1033, te: 1284-1356) and the intercept appended by
    ``GAM._validate_data_dep_params`` (pygam.py:316-317)."""
    out = []
    for d in term_specs:
        k = d["kind"]
        if k == "s":
            out.append(_norm_spline(d))
        elif k == "l":
            out.append(_norm_linear(d))
        elif k == "f":
            pens = _as_list(d.get("penalties", "auto"))
            lam = _as_list(d.get("lam", 0.6))
            if len(lam) == 1:
                lam = lam * len(pens)
            out.append(dict(kind="f", feature=int(d["feature"]), penalties=pens,
                            lam=[float(v) for v in lam], constraints=[None], dtype="categorical",
                            coding=d.get("coding", "one-hot"), spline_order=0, basis="ps",
                            by=None, n_splines=None, edge_knots=None))
        elif k == "te":
            # marginals are terms of their own: s() (default n_splines 10) or l() (terms.py:1330-1356; tensors of tensors
            # are refused there, factor marginals are not restated here)
            margs = [_norm_linear(m) if m.get("kind") == "l" else _norm_spline(m, default_n_splines=10)
                     for m in d["marginals"]]
            if len(margs) < 2:
                raise ValueError("TensorTerm requires at least 2 marginal terms")
            out.append(dict(kind="te", marginals=margs, by=d.get("by")))
        elif k == "intercept":
            out.append(dict(kind="intercept"))
        else:
            raise ValueError(f"unknown term kind {k}")
    if fit_intercept and not any(t["kind"] == "intercept" for t in out):
        out.append(dict(kind="intercept"))
    if not out:
        raise ValueError("At least 1 term must be specified")
    return out


def n_coefs(t):
    """terms.py:563 (intercept), 678 (linear), 893 (spline), 1101 (factor), 1428 (tensor)"""
    k = t["kind"]
    if k in ("intercept", "l"):
        return 1
    if k == "s":
        return t["n_splines"]
    if k == "f":
        return t["n_splines"] - (1 if t["coding"] == "dummy" else 0)
    if k == "te":
        return int(np.prod([n_coefs(m) for m in t["marginals"]]))
    raise ValueError(k)


def total_coefs(terms):
    return sum(n_coefs(t) for t in terms)


def coef_slices(terms):
    """terms.py:1874-1899"""
    out, start = [], 0
    for t in terms:
        out.append(slice(start, start + n_coefs(t)))
        start += n_coefs(t)
    return out


def has_constraint(terms):
    """terms.py:260-263, 1418-1423, 1861-1867 -- note the string "none" counts as a constraint"""
    for t in terms:
        if t["kind"] == "te":
            if any(c is not None for m in t["marginals"] for c in m["constraints"]):
                return True
        elif t["kind"] != "intercept":
            if any(c is not None for c in t["constraints"]):
                return True
    return False


def gen_edge_knots(col, dtype):
    """utils.py:592-621"""
    if dtype == "categorical":
        return np.r_[np.min(col) - 0.5, np.max(col) + 0.5]
    return np.r_[np.min(col), np.max(col)]


def compile_terms(terms, X):
    """data-dependent parameters: terms.py:668-691 (linear), 895-924 (spline keeps user
    edge knots), 1054-1075 (factor recomputes), 1430-1452 (tensor), 1810-1833 (list)"""
    terms = copy.deepcopy(terms)
    for t in terms:
        k = t["kind"]
        if k == "l":
            t["edge_knots"] = gen_edge_knots(X[:, t["feature"]], "numerical")
        elif k == "s":
            if t["edge_knots"] is None:
                t["edge_knots"] = gen_edge_knots(X[:, t["feature"]], t["dtype"])
        elif k == "f":
            t["n_splines"] = len(np.unique(X[:, t["feature"]]))
            t["edge_knots"] = gen_edge_knots(X[:, t["feature"]], "categorical")
        elif k == "te":
            for m in t["marginals"]:
                if m["edge_knots"] is None:
                    m["edge_knots"] = gen_edge_knots(X[:, m["feature"]], m["dtype"])  # l(): terms.py:668-691
    return terms


# ======================================================================================
# basis (utils.py:624-770) and tensor product (utils.py:899-948)
# ======================================================================================
def bspline_basis(x, edge_knots, n_splines=20, spline_order=3, periodic=False):
    """Dense (len(x), n_splines) B-spline design: uniform knots on the normalised range,
    Cox-de Boor recursion evaluated for all columns at once, cyclic fold, and linear
    extrapolation outside the edge knots.  Arithmetic follows utils.py:678-765 step by step
    (same knot construction incl. the 1e-9 nudges) so values agree to the last bit."""
    x = np.ravel(np.asarray(x, dtype=np.float64))
    K = n_splines + spline_order * bool(periodic)  # utils.py:678
    ek = np.sort(np.asarray(edge_knots, dtype=np.float64))
    lo = ek[0]
    span = ek[-1] - ek[0]
    if span == 0:
        span = 1
    interior = np.linspace(0, 1, 1 + K - spline_order)  # utils.py:686
    h = np.diff(interior[:2])[0]
    u = (x - lo) / span  # utils.py:690
    if periodic:
        u = u % (1 + 1e-9)  # utils.py:694
    u = np.r_[u, 0.0, 1.0]  # two sentinel rows: values at the edges (utils.py:697)
    below, above = u < 0, u > 1
    inside = ~(above + below)
    ucol = u[:, None]

    pad = np.arange(1, spline_order + 1) * h  # utils.py:708-710
    knots = np.r_[-pad[::-1], interior, 1 + pad]
    knots[-1] += 1e-9

    # order-0 indicator functions (utils.py:713-714)
    B = (ucol >= knots[:-1]).astype(int) * (ucol < knots[1:]).astype(int)
    B[-1] = B[-2][::-1]

    width = len(knots) - 1
    edge_prev = None
    for m in range(2, spline_order + 2):  # utils.py:717-734
        width -= 1
        a = ucol - knots[:width]
        a = a * B[:, :width]
        left = a / (knots[m - 1: width + m - 1] - knots[:width])
        b = (knots[m: width + m] - ucol) * B[:, 1: width + 1]
        right = b / (knots[m: width + m] - knots[1: width + 1])
        edge_prev = B[-2:]
        B = left + right

    if periodic and spline_order > 0:  # utils.py:736-742
        B[:, :spline_order] = np.max([B[:, :spline_order], B[:, -spline_order:]], axis=0)
        B = B[:, :-spline_order]

    if (above.any() or below.any()) and spline_order > 0:  # utils.py:747-763
        B[~inside] = 0.0
        gl = edge_prev[:, :-1] / (knots[spline_order:-1] - knots[: -spline_order - 1])
        gr = edge_prev[:, 1:] / (knots[spline_order + 1:] - knots[1:-spline_order])
        slope = spline_order * (gl - gr)
        if below.any():
            B[below] = slope[0] * ucol[below] + B[-2]
        if above.any():
            B[above] = slope[1] * (ucol[above] - 1) + B[-1]
    return np.asarray(B[:-2], dtype=np.float64)


def tensor_rows(a, b):
    """row-wise Kronecker product, first factor slowest (utils.py:943-946)"""
    return (a[:, :, None] * b[:, None, :]).reshape(a.shape[0], a.shape[1] * b.shape[1])


def _spline_columns(t, X):
    """terms.py:926-956"""
    B = bspline_basis(X[:, t["feature"]], t["edge_knots"], n_splines=t["n_splines"],
                      spline_order=t["spline_order"], periodic=t["basis"] == "cp")
    if t.get("by") is not None:
        B = B * X[:, t["by"]][:, None]
    return B


def term_columns(t, X):
    k = t["kind"]
    n = X.shape[0]
    if k == "intercept":  # terms.py:584-599
        return np.ones((n, 1))
    if k == "l":  # terms.py:693-708
        return X[:, t["feature"]][:, None].astype(np.float64)
    if k == "s":
        return _spline_columns(t, X)
    if k == "f":  # terms.py:1077-1096
        B = _spline_columns(t, X)
        return B[:, 1:] if t["coding"] == "dummy" else B
    if k == "te":  # terms.py:1454-1477
        B = term_columns(t["marginals"][0], X)  # every marginal builds its own columns (terms.py:1469-1472)
        for m in t["marginals"][1:]:
            B = tensor_rows(B, term_columns(m, X))
        if t.get("by") is not None:
            B = B * X[:, t["by"]][:, None]
        return B
    raise ValueError(k)


def model_matrix(terms, X):
    """terms.py:1901-1923 -- sparse CSC (n, m)"""
    X = np.asarray(X, dtype=np.float64)
    return sps.hstack([sps.csc_array(term_columns(t, X)) for t in terms], format="csc")


# ======================================================================================
# penalties and constraints (penalties.py)
# ======================================================================================
def _diff_matrix(n, order):
    """(n, n-order) matrix whose columns are order-th differences: the ``D`` of
    penalties.py:27-29 / 312-348 (``sparse_diff`` along the last axis of the identity)"""
    D = np.eye(n)
    for _ in range(order):
        D = D[:, 1:] - D[:, :-1]
    return D


def pen_derivative(n, coef=None, derivative=2, periodic=False):
    """penalties.py:9-49"""
    if n == 1:
        return np.zeros((1, 1))
    D = _diff_matrix(n + 2 * derivative * periodic, derivative)
    if periodic:
        head = D[:, :derivative].copy()
        D[:, -2 * derivative: -derivative] += head * (-1) ** derivative
        half = int((n + 2 * derivative) / 2)
        D[-half:] = D[:half][::-1, ::-1]
        D = D[derivative:-derivative, derivative:-derivative]
    return D @ D.T


def pen_l2(n, coef=None):
    """penalties.py:56-73"""
    return np.eye(n)


def pen_none(n, coef=None):
    """penalties.py:261-276"""
    return np.zeros((n, n))


def pen_periodic(n, coef=None):
    """penalties.py:52-53"""
    return pen_derivative(n, coef, derivative=2, periodic=True)


def con_monotonic(n, coef, increasing):
    """penalties.py:76-113"""
    coef = np.ravel(coef)
    if n != len(coef):
        raise ValueError("dimension mismatch")
    if n == 1:
        return np.zeros((1, 1))
    d = np.diff(coef)
    viol = (d < 0) if increasing else (d > 0)
    D = _diff_matrix(n, 1) * viol.astype(float)[None, :]
    return D @ D.T


def con_convexity(n, coef, convex):
    """penalties.py:153-188"""
    coef = np.ravel(coef)
    if n != len(coef):
        raise ValueError("dimension mismatch")
    if n == 1:
        return np.zeros((1, 1))
    d = np.diff(coef, n=2)
    viol = (d < 0) if convex else (d > 0)
    D = _diff_matrix(n, 2) * viol.astype(float)[None, :]
    return D @ D.T


PENALTIES = {"derivative": pen_derivative, "l2": pen_l2, "none": pen_none, "periodic": pen_periodic}
CONSTRAINTS = {
    "monotonic_inc": lambda n, c: con_monotonic(n, c, True),
    "monotonic_dec": lambda n, c: con_monotonic(n, c, False),
    "convex": lambda n, c: con_convexity(n, c, True),
    "concave": lambda n, c: con_convexity(n, c, False),
    "none": lambda n, c: np.zeros((n, n)),
}


def _simple_term_penalty(t):
    """terms.py:307-349"""
    if t["kind"] == "intercept":
        return np.array([[0.0]])
    n = n_coefs(t)
    total = 0
    for pen, lam in zip(t["penalties"], t["lam"]):
        if pen == "auto":
            if t["dtype"] == "numerical":
                if t["kind"] == "s":
                    pen = "periodic" if t["basis"] == "cp" else "derivative"
                else:
                    pen = "l2"
            if t["dtype"] == "categorical":
                pen = "l2"
        if pen is None:
            pen = "none"
        fn = PENALTIES[pen] if isinstance(pen, str) else pen
        total = total + np.multiply(np.asarray(fn(n, None)), lam)
    return np.asarray(total, dtype=np.float64)


def term_penalty(t):
    if t["kind"] != "te":
        return _simple_term_penalty(t)
    # terms.py:1479-1517: sum_i I (x) ... (x) P_i (x) ... (x) I
    total = 0
    for i in range(len(t["marginals"])):
        acc = None
        for j, m in enumerate(t["marginals"]):
            Pj = _simple_term_penalty(m) if j == i else np.eye(n_coefs(m))
            acc = Pj if acc is None else np.kron(acc, Pj)
        total = total + acc
    return total


def penalty_matrix(terms):
    """terms.py:1925-1947 -- dense (m, m)"""
    return scipy.linalg.block_diag(*[term_penalty(t) for t in terms])


def _simple_term_constraint(t, coef, clam, cl2):
    """terms.py:351-396"""
    if t["kind"] == "intercept":
        return np.array([[0.0]])
    n = n_coefs(t)
    total = 0
    for c in t["constraints"]:
        if c is None:
            c = "none"
        fn = CONSTRAINTS[c] if isinstance(c, str) else c
        total = total + np.asarray(fn(n, coef)) * clam
    total = np.asarray(total, dtype=np.float64)
    if np.count_nonzero(total) > 0:  # `Cs.nnz > 0` on a sparse matrix
        total = total + cl2 * np.eye(n)
    return total


def term_constraint(t, coef, clam, cl2):
    if t["kind"] != "te":
        return _simple_term_constraint(t, coef, clam, cl2)
    # terms.py:1519-1641: every 1-d slice of the coefficient tensor along marginal i
    dims = [n_coefs(m) for m in t["marginals"]]
    size = int(np.prod(dims))
    C = np.zeros((size, size))
    idx = np.arange(size).reshape(dims)
    for i, m in enumerate(t["marginals"]):
        rows = np.moveaxis(idx, i, 0).reshape(dims[i], size // dims[i])
        for sl in rows.T:
            C[np.ix_(sl, sl)] += _simple_term_constraint(m, coef[sl], clam, cl2)
    return C


def constraint_matrix(terms, coef, clam=CONSTRAINT_LAM, cl2=CONSTRAINT_L2):
    """terms.py:1949-1979"""
    return scipy.linalg.block_diag(
        *[term_constraint(t, coef[sl], clam, cl2) for t, sl in zip(terms, coef_slices(terms))])


# ======================================================================================
# links (links.py) and distributions (distributions.py)
# ======================================================================================
def link_fn(name, mu, levels=1):
    if name == "identity":
        return mu
    if name == "logit":
        return np.log(mu) - np.log(levels - mu)  # links.py:105
    if name == "log":
        return np.log(mu)
    if name == "inverse":
        return mu ** -1.0
    if name == "inv_squared":
        return mu ** -2.0
    raise ValueError(name)


def link_inv(name, lp, levels=1):
    if name == "identity":
        return lp
    if name == "logit":
        e = np.exp(lp)  # links.py:121-122 (overflows to nan for lp > 709, as the reference)
        return levels * e / (e + 1)
    if name == "log":
        return np.exp(lp)
    if name == "inverse":
        return lp ** -1.0
    if name == "inv_squared":
        return lp ** -0.5
    raise ValueError(name)


def link_grad(name, mu, levels=1):
    if name == "identity":
        return np.ones_like(mu)
    if name == "logit":
        return levels / (mu * (levels - mu))  # links.py:137
    if name == "log":
        return 1.0 / mu
    if name == "inverse":
        return -1 * mu ** -2.0
    if name == "inv_squared":
        return -2 * mu ** -3.0
    raise ValueError(name)


def variance(dist, mu, levels=1):
    """distributions.py:160, 265, 370, 467, 570 (called without weights at pygam.py:632)"""
    if dist == "normal":
        return np.ones_like(mu)
    if dist == "binomial":
        return mu * (1 - mu / levels)
    if dist == "poisson":
        return mu
    if dist == "gamma":
        return mu ** 2
    if dist == "inv_gauss":
        return mu ** 3
    raise ValueError(dist)


def ylogydu(y, u):
    """utils.py:773-789"""
    mask = np.atleast_1d(y) != 0.0
    out = np.zeros_like(u)
    out[mask] = y[mask] * np.log(y[mask] / u[mask])
    return out


def deviance_rows(dist, y, mu, levels=1, scale=None, scaled=True, weights=None):
    """distributions.py:162-185, 267-291, 372-397, 469-494, 572-597 + decorator 13-20"""
    if dist == "normal":
        dev = (y - mu) ** 2
        if scaled:
            dev = dev / scale ** 2
    else:
        if dist == "binomial":
            dev = 2 * (ylogydu(y, mu) + ylogydu(levels - y, levels - mu))
        elif dist == "poisson":
            dev = 2 * (ylogydu(y, mu) - (y - mu))
        elif dist == "gamma":
            dev = 2 * ((y - mu) / mu - np.log(y / mu))
        elif dist == "inv_gauss":
            dev = ((y - mu) ** 2) / (mu ** 2 * y)
        else:
            raise ValueError(dist)
        if scaled:
            dev = dev / scale
    if weights is None:
        weights = np.ones_like(mu)
    return dev * weights


def log_pdf_rows(dist, y, mu, weights, levels=1, scale=None):
    """distributions.py:110-131, 225-247, 323-352, 428-449, 531-552"""
    if dist == "normal":
        return scipy.stats.norm.logpdf(y, loc=mu, scale=scale / weights)
    if dist == "binomial":
        return scipy.stats.binom.logpmf(y, levels, mu / levels)
    if dist == "poisson":
        return scipy.stats.poisson.logpmf(y, mu=mu * weights)
    if dist == "gamma":
        nu = weights / scale
        return scipy.stats.gamma.logpdf(x=y, a=nu, scale=mu / nu)
    if dist == "inv_gauss":
        return scipy.stats.invgauss.logpdf(y, mu, scale=1.0 / (weights / scale))
    raise ValueError(dist)


# ======================================================================================
# model families (pygam.py:2371-3551)
# ======================================================================================
FAMILIES = {
    # model name -> (distribution, link, scale known?)
    "LinearGAM": ("normal", "identity"),
    "LogisticGAM": ("binomial", "logit"),
    "PoissonGAM": ("poisson", "log"),
    "GammaGAM": ("gamma", "log"),
    "InvGaussGAM": ("inv_gauss", "log"),
    "ExpectileGAM": ("normal", "identity"),
}


def family_of(model, model_kwargs=None):
    kw = dict(model_kwargs or {})
    if model == "GAM":
        dist, link = kw.get("distribution", "normal"), kw.get("link", "identity")
    else:
        dist, link = FAMILIES[model]
    fam = dict(model=model, dist=dist, link=link, levels=1,
               expectile=kw.get("expectile", 0.5) if model == "ExpectileGAM" else None)
    # known scale: binomial / poisson have scale 1 (distributions.py:222, 320); the others
    # only if the user passes one (pygam.py:2447-2450)
    if dist in ("binomial", "poisson"):
        fam["scale"], fam["known_scale"] = 1.0, True
    else:
        fam["scale"] = kw.get("scale")
        fam["known_scale"] = fam["scale"] is not None
    fam["max_iter"] = kw.get("max_iter", 100)
    fam["tol"] = kw.get("tol", 1e-4)
    fam["fit_intercept"] = kw.get("fit_intercept", True)
    return fam


def prepare_weights(model, y, weights=None, exposure=None):
    """float32 casts of pygam.py:888-895 and the Poisson exposure handling 2838-2890"""
    y = np.asarray(y, dtype=np.float64).ravel()
    if model == "PoissonGAM":
        if exposure is not None:
            exposure = np.array(exposure).astype("f").ravel()
        else:
            exposure = np.ones_like(y).astype("float64")
        y = y / exposure
        if weights is not None:
            weights = np.array(weights).astype("f").ravel()
        else:
            weights = np.ones_like(y).astype("float64")
        weights = weights * exposure
    if weights is not None:
        weights = np.array(weights).astype("f").ravel()
    else:
        weights = np.ones_like(y).astype("float64")
    return y, weights


# ======================================================================================
# PIRLS (pygam.py:499-822)
# ======================================================================================
def dense_cholesky_upper(A):
    """utils.py:77-91"""
    try:
        return scipy.linalg.cholesky(A, lower=False)
    except scipy.linalg.LinAlgError as err:
        raise NotPositiveDefiniteError("Matrix is not positive definite") from err


def cholesky_with_retry(A, state):
    """pygam.py:499-537 incl. the literal subtract/add of the diagonal loading"""
    A = np.array(A, dtype=np.float64)
    eye = np.eye(A.shape[0])
    l2 = state["constraint_l2"]
    while l2 <= CONSTRAINT_L2_MAX:
        try:
            L = dense_cholesky_upper(A)
            state["constraint_l2"] = l2
            return L
        except NotPositiveDefiniteError:
            A -= l2 * eye
            l2 *= 10
            A += l2 * eye
    raise NotPositiveDefiniteError("Matrix is not positive \ndefinite.")


def irls_weight(fam, mu, weights, y):
    """square-root IRLS weight, pygam.py:599-636 and ExpectileGAM 3419-3460"""
    g = link_grad(fam["link"], mu, fam["levels"])
    w = (g ** 2 * variance(fam["dist"], mu, fam["levels"]) * weights ** -1) ** -0.5
    if fam["expectile"] is not None:
        asym = (y > mu) * fam["expectile"] + (y <= mu) * (1 - fam["expectile"])
        w = w * asym ** 0.5
    return w


def weight_mask(w):
    """pygam.py:638-663"""
    mask = (np.abs(w) >= np.sqrt(EPS)) * np.isfinite(w)
    if mask.sum() == 0:
        raise OptimizationError("PIRLS optimization has diverged.")
    return mask


def initial_estimate(fam, y, modelmat):
    """pygam.py:665-710"""
    m = modelmat.shape[1]
    if fam["model"] == "LinearGAM":
        return np.ones(m) * np.sqrt(EPS)
    y = np.array(y, dtype=np.float64)
    y[y == 0] += 0.01
    y[y == 1] -= 0.01
    y_ = link_fn(fam["link"], y, fam["levels"])[:, None]
    assert np.isfinite(y_).all()
    BtB = modelmat.T.dot(modelmat).toarray()
    BtB = BtB + np.eye(m) * np.sqrt(EPS)  # load_diagonal, utils.py:451-474
    return np.linalg.solve(BtB, modelmat.T.dot(y_)).ravel()


def pirls(fam, terms, X, y, weights, coef=None, timer=None):
    """One full PIRLS fit on prepared inputs; returns coefficient, per-iteration logs and
    the last-iteration factors needed by the statistics (pygam.py:715-822)."""
    modelmat = model_matrix(terms, X)
    n, m = modelmat.shape
    if coef is None or len(coef) != m or not np.isfinite(coef).all():
        coef = initial_estimate(fam, y, modelmat)
    coef = np.asarray(coef, dtype=np.float64)
    assert np.isfinite(coef).all()
    state = dict(constraint_l2=CONSTRAINT_L2)
    P = penalty_matrix(terms)
    S = np.eye(m) * np.sqrt(EPS)
    constrained = has_constraint(terms)
    if not constrained:
        E = cholesky_with_retry(S + P, state)
    logs = dict(deviance=[], diffs=[], accuracy=[], time=[])
    diff = np.inf
    for _ in range(fam["max_iter"]):
        if constrained:
            P = penalty_matrix(terms)
            C = constraint_matrix(terms, coef, CONSTRAINT_LAM, state["constraint_l2"])
            E = cholesky_with_retry(S + P + C, state)
        lp = modelmat.dot(coef).ravel()
        mu = link_inv(fam["link"], lp, fam["levels"])
        w = irls_weight(fam, mu, weights, y)
        mask = weight_mask(w)
        ym, lpm, mum, wm = y[mask], lp[mask], mu[mask], w[mask]
        pseudo = wm * (lpm + (ym - mum) * link_grad(fam["link"], mum, fam["levels"]))
        # on_loop_start callbacks (callbacks.py:125-141, 159-174)
        if timer is not None:
            logs["time"].append(timer())
        logs["deviance"].append(
            deviance_rows(fam["dist"], ym, mum, fam["levels"], fam["scale"], scaled=False).sum())
        logs["accuracy"].append(np.mean(ym == (mum > 0.5)))
        WB = sps.diags(wm).dot(modelmat[mask, :])
        Q, R = np.linalg.qr(WB.toarray())
        if not np.isfinite(Q).all() or not np.isfinite(R).all():
            raise ValueError("QR decomposition produced NaN or Inf. Check X data.")
        r = int(np.min([m, n, mask.sum()]))
        U, d, Vt = np.linalg.svd(np.vstack([R, E]), full_matrices=False)
        d_inv = d ** -1
        U1 = U[:r, :r]
        Vt = Vt[:r]
        d_inv = d_inv[:r]
        Bmat = (Vt.T * d_inv).dot(U1.T).dot(Q.T)
        coef_new = Bmat.dot(pseudo).ravel()
        diff = np.linalg.norm(coef - coef_new) / np.linalg.norm(coef_new)
        coef = coef_new
        logs["diffs"].append(diff)
        if diff < fam["tol"]:
            break
    return dict(coef=coef, logs=logs, modelmat=modelmat, B=Bmat, U1=U1, converged=bool(diff < fam["tol"]),
                constraint_l2=state["constraint_l2"])


def model_statistics(fam, res, y, weights):
    """pygam.py:993-1228; mutates fam['scale'] when it is estimated, like the reference
    mutates ``distribution.scale``"""
    modelmat, coef, Bmat, U1 = res["modelmat"], res["coef"], res["B"], res["U1"]
    n = len(y)
    lp = modelmat.dot(coef).ravel()
    mu = link_inv(fam["link"], lp, fam["levels"])
    st = {}
    st["edof_per_coef"] = np.diagonal(U1.dot(U1.T))
    st["edof"] = st["edof_per_coef"].sum()
    if not fam["known_scale"]:
        # distributions.py:78
        phi = np.sum(weights * variance(fam["dist"], mu, fam["levels"]) ** -1 * (y - mu) ** 2) / (n - st["edof"])
        fam["scale"] = phi ** 0.5
    st["scale"] = fam["scale"]
    st["cov"] = Bmat.dot(Bmat.T) * fam["scale"] ** 2
    st["se"] = st["cov"].diagonal() ** 0.5

    def loglik(y_, mu_):
        yy = y_
        if fam["model"] == "PoissonGAM":  # pygam.py:2797-2798
            yy = np.round(y_ * weights).astype("int")
        return log_pdf_rows(fam["dist"], yy, mu_, weights, fam["levels"], fam["scale"]).sum()

    ll = loglik(y, mu)
    st["loglikelihood"] = ll
    st["AIC"] = -2 * ll + 2 * st["edof"] + 2 * (not fam["known_scale"])  # pygam.py:1077-1084
    st["AICc"] = st["AIC"] + 2 * (st["edof"] + 1) * (st["edof"] + 2) / (n - st["edof"] - 2)
    # pseudo r2, pygam.py:1141-1154
    null_mu = y.mean() * np.ones_like(y).astype("float64")
    null_d = deviance_rows(fam["dist"], y, null_mu, fam["levels"], fam["scale"], True, weights)
    full_d = deviance_rows(fam["dist"], y, mu, fam["levels"], fam["scale"], True, weights)
    null_ll = loglik(y, null_mu)
    st["pseudo_r2"] = dict(explained_deviance=1.0 - full_d.sum() / null_d.sum(),
                           McFadden=ll / null_ll, McFadden_adj=1.0 - (ll - st["edof"]) / null_ll)
    # GCV / UBRE, pygam.py:1207-1228 (gamma = 1.4, `~add_scale` == -2)
    dev = deviance_rows(fam["dist"], y, mu, fam["levels"], fam["scale"], False, weights).sum()
    gamma = 1.4
    st["GCV"], st["UBRE"] = None, None
    if fam["known_scale"]:
        st["UBRE"] = 1.0 / n * dev - (-2) * fam["scale"] + 2.0 * gamma / n * st["edof"] * fam["scale"]
    else:
        st["GCV"] = (n * dev) / (n - gamma * st["edof"]) ** 2
    st["deviance"] = full_d.sum()
    st["n_samples"] = n
    return st


def fit(case, X, y, weights=None, exposure=None, coef=None, terms=None, timer=None):
    """``GAM.fit`` (pygam.py:860-915) for a case dict of ``tests/golden/cases.py``"""
    fam = family_of(case["model"], case.get("model_kwargs"))
    X = np.asarray(X, dtype=np.float64)
    y, weights = prepare_weights(case["model"], y, weights, exposure)
    if terms is None:
        terms = compile_terms(normalize_terms(case["terms"], fam["fit_intercept"]), X)
    res = pirls(fam, terms, X, y, weights, coef=coef, timer=timer)
    res["statistics"] = model_statistics(fam, res, y, weights)
    res["terms"], res["fam"] = terms, fam
    return res


def predict_mu(res, X):
    """pygam.py:423-450"""
    lp = model_matrix(res["terms"], np.asarray(X, dtype=np.float64)).dot(res["coef"]).ravel()
    return link_inv(res["fam"]["link"], lp, res["fam"]["levels"])


# ======================================================================================
# grid search (pygam.py:1808-2085)
# ======================================================================================
def _cartesian(lists):
    """utils.py:792-816"""
    out = [[]]
    for lst in lists:
        out = [prev + [v] for prev in out for v in lst]
    return out


def set_lams(terms, values):
    """plural ``lam`` assignment: terms.py:449-480 (values consumed in term order, each term
    taking as many as it has penalties; scalar broadcasts)"""
    terms = copy.deepcopy(terms)
    slots = []
    for t in terms:
        if t["kind"] == "intercept":
            continue
        if t["kind"] == "te":
            for m in t["marginals"]:
                slots.append(m)
        else:
            slots.append(t)
    size = sum(len(s["lam"]) for s in slots)
    if np.ndim(values) == 0:
        values = [float(values)] * size
    values = [float(v) for v in np.ravel(values)]
    if len(values) != size:
        raise ValueError(f"Expected lam to have length {size}")
    pos = 0
    for s in slots:
        k = len(s["lam"])
        s["lam"] = values[pos: pos + k]
        pos += k
    return terms


def gridsearch(case, X, y, weights=None, exposure=None, lam_grid=None, objective="auto"):
    """Sequential warm-started refits over a lam grid, best by GCV/UBRE.
    ``lam_grid``: None (default 11-point logspace, pygam.py:1965-1966), a 1-d array (same lam
    on every term), a list of per-term lists (cartesian product, pygam.py:2001-2004) or a 2-d
    ndarray (rows as given, pygam.py:1991)."""
    fam0 = family_of(case["model"], case.get("model_kwargs"))
    X = np.asarray(X, dtype=np.float64)
    terms = compile_terms(normalize_terms(case["terms"], fam0["fit_intercept"]), X)
    if objective == "auto":
        objective = "UBRE" if fam0["known_scale"] else "GCV"
    if lam_grid is None:
        lam_grid = np.logspace(-3, 3, 11)
    if isinstance(lam_grid, np.ndarray) and lam_grid.ndim == 2:
        candidates = [list(row) for row in lam_grid]
    elif any(np.ndim(g) > 0 for g in lam_grid):
        candidates = _cartesian([list(np.atleast_1d(g)) for g in lam_grid])
    else:
        candidates = [float(v) for v in lam_grid]
    results, scores = [], []
    coef = None
    for cand in candidates:
        try:
            res = fit(case, X, y, weights=weights, exposure=exposure, coef=coef,
                      terms=set_lams(terms, cand))
        except ValueError:
            continue
        coef = res["coef"]
        results.append(res)
        scores.append(res["statistics"][objective])
    best = int(np.argmin(scores))
    # `scores[-1] < best_score` keeps the first of equal scores, as argmin does
    return dict(results=results, scores=np.asarray(scores), best=best, objective=objective,
                candidates=candidates)
