"""Declarative parity cases shared by the golden generator, the oracle tests and the GPU parity tests.

Every case is a plain dict so that the same description can be turned into
(a) reference pyGAM objects (``make_golden.py``, run only in the authoring container),
(b) an oracle model spec (``oracle/gam_oracle.py``) and
(c) ``pygam_b200`` model objects (the product).

Term dicts use the reference's own keyword names (``/root/reference/pygam/terms.py:810-823``
for ``s``, ``1010-1012`` for ``f``, ``662-666`` for ``l``, ``1284-1319`` for ``te``).
The BASELINE.json configs are reproduced at reduced row counts (C1 is full size).
"""

import numpy as np


def s(feature, **kw):
    d = dict(kind="s", feature=feature)
    d.update(kw)
    return d


def l(feature, **kw):  # noqa: E743
    d = dict(kind="l", feature=feature)
    d.update(kw)
    return d


def f(feature, **kw):
    d = dict(kind="f", feature=feature)
    d.update(kw)
    return d


def te(*marginals, **kw):
    d = dict(kind="te", marginals=list(marginals))
    d.update(kw)
    return d


# --------------------------------------------------------------------------------------
# synthetic data; numpy Generator streams are stable across versions
# for uniform/normal draws, and the drawn arrays are stored in the fixtures anyway.
# --------------------------------------------------------------------------------------
def data_c2(n, seed=1234):
    rng = np.random.default_rng(seed)
    X = rng.random((n, 10))
    eta = 0.5 * sum(np.sin(2 * np.pi * (j + 1) * X[:, j] / 3.0) for j in range(10))
    y = (rng.random(n) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
    return X, y


def data_c3(n, seed=1234):
    rng = np.random.default_rng(seed)
    X = rng.random((n, 24))
    eta = 0.5 + 0.075 * np.sin(2 * np.pi * X[:, :20]).sum(1)
    eta = eta + 0.5 * X[:, 20] * X[:, 21] - 0.5 * X[:, 22] * X[:, 23]
    y = rng.poisson(np.exp(eta)).astype(np.float64)
    return X, y


def data_c4(n, seed=1234):
    rng = np.random.default_rng(seed)
    X = rng.random((n, 3))
    eta = np.sin(2 * np.pi * X[:, 0]) + X[:, 1] ** 2 + 0.5 * np.cos(4 * np.pi * X[:, 2])
    y = eta + 0.3 * rng.standard_normal(n)
    return X, y


def data_c5(n, seed=1234):
    rng = np.random.default_rng(seed)
    X = rng.random((n, 4))
    eta = np.sqrt(X[:, 0]) + np.sin(3 * X[:, 1]) + X[:, 2] ** 2 + 0.2 * np.cos(6 * X[:, 3])
    y = eta + 0.3 * rng.standard_normal(n)
    return X, y


def data_positive(n, seed=7):
    """positive responses for the gamma / inverse-gaussian families"""
    rng = np.random.default_rng(seed)
    X = rng.random((n, 3))
    eta = 0.3 + 0.8 * np.sin(2 * np.pi * X[:, 0]) * 0.5 + X[:, 1] - 0.5 * X[:, 2] ** 2
    mu = np.exp(eta)
    y = rng.gamma(shape=5.0, scale=mu / 5.0)
    return X, y


def data_mixed(n, seed=11):
    """numerical + categorical + by-variable columns"""
    rng = np.random.default_rng(seed)
    X = np.empty((n, 5))
    X[:, 0] = rng.random(n) * 10 - 5
    X[:, 1] = rng.random(n)
    X[:, 2] = rng.integers(0, 6, n)
    X[:, 3] = rng.standard_normal(n)
    X[:, 4] = rng.random(n) * 2 * np.pi
    eta = np.sin(X[:, 0]) + 2 * X[:, 1] + 0.3 * X[:, 2] + 0.5 * X[:, 3] * X[:, 1] + np.cos(X[:, 4])
    y = eta + 0.2 * rng.standard_normal(n)
    return X, y


def data_counts_exposure(n, seed=5):
    rng = np.random.default_rng(seed)
    X = rng.random((n, 2))
    exposure = rng.integers(1, 6, n).astype(np.float64)
    rate = np.exp(0.2 + np.sin(2 * np.pi * X[:, 0]) + 0.5 * X[:, 1])
    y = rng.poisson(rate * exposure).astype(np.float64)
    return X, y, exposure


# --------------------------------------------------------------------------------------
# fit cases
# --------------------------------------------------------------------------------------
def _c2_terms():
    return [s(j, n_splines=25) for j in range(10)]


def _c3_terms():
    return [s(j) for j in range(20)] + [te(s(20, n_splines=10), s(21, n_splines=10)),
                                         te(s(22, n_splines=10), s(23, n_splines=10))]


def _c5_terms():
    return [s(0, constraints="monotonic_inc"), s(1), s(2, constraints="monotonic_inc"),
            s(3, constraints="convex")]


FIT_CASES = {
    # BASELINE config 1 (full size)
    "c1_wage": dict(model="LinearGAM", terms=[s(0), s(1), f(2)], data=("wage",)),
    # BASELINE config 2 at n_ref rows
    "c2_logistic": dict(model="LogisticGAM", terms=_c2_terms(), data=("c2", 3000)),
    # BASELINE config 3 at n_ref rows
    "c3_poisson": dict(model="PoissonGAM", terms=_c3_terms(), data=("c3", 4000)),
    # BASELINE config 4 model (single fit; the grid is in GRID_CASES)
    "c4_linear": dict(model="LinearGAM", terms=[s(0), s(1), s(2)], data=("c4", 3000)),
    # BASELINE config 5 at n_ref rows
    "c5_expectile": dict(model="ExpectileGAM", terms=_c5_terms(), data=("c5", 3000),
                         model_kwargs=dict(expectile=0.9)),
    "expectile_plain": dict(model="ExpectileGAM", terms=[s(0), s(1)], data=("c5", 1500),
                            model_kwargs=dict(expectile=0.25)),
    "gamma": dict(model="GammaGAM", terms=[s(0), s(1), s(2)], data=("positive", 2000)),
    "invgauss": dict(model="InvGaussGAM", terms=[s(0), s(1), s(2)], data=("positive", 2000)),
    "linear_weights": dict(model="LinearGAM", terms=[s(0), s(1), s(2)], data=("c4", 2000),
                           weights="uniform"),
    "logistic_weights": dict(model="LogisticGAM", terms=[s(0), s(1, n_splines=12)],
                             data=("c2", 2000), weights="uniform"),
    "poisson_exposure": dict(model="PoissonGAM", terms=[s(0), s(1)],
                             data=("counts_exposure", 2000)),
    "mixed_terms": dict(model="LinearGAM",
                        terms=[s(0, n_splines=12, spline_order=2), l(1), f(2),
                               s(3, by=1, n_splines=8), s(4, basis="cp", n_splines=10)],
                        data=("mixed", 2500)),
    "dummy_coding": dict(model="LinearGAM", terms=[s(0), f(2, coding="dummy")],
                         data=("mixed", 1500)),
    "orders": dict(model="LinearGAM",
                   terms=[s(0, n_splines=9, spline_order=1), s(1, n_splines=7, spline_order=4),
                          s(3, n_splines=6, spline_order=0), s(4, n_splines=5, spline_order=3)],
                   data=("mixed", 1500)),
    "tensor3": dict(model="LinearGAM",
                    terms=[te(s(0, n_splines=5), s(1, n_splines=4), s(4, n_splines=6, spline_order=2)),
                           s(3)],
                    data=("mixed", 2000)),
    "tensor_by": dict(model="PoissonGAM",
                      terms=[te(s(0, n_splines=6), s(1, n_splines=6), by=1), s(0)],
                      data=("counts_exposure2", 1500)),
    "multi_penalty": dict(model="LinearGAM",
                          terms=[s(0, penalties=["derivative", "l2"], lam=[0.6, 0.1]),
                                 s(1, penalties=None), l(2, lam=3.0)],
                          data=("c4", 1500)),
    "concave_dec": dict(model="LinearGAM",
                        terms=[s(0, constraints="monotonic_dec"), s(1, constraints="concave")],
                        data=("c5", 1500)),
    "no_intercept": dict(model="LinearGAM", terms=[s(0), s(1)], data=("c4", 1200),
                         model_kwargs=dict(fit_intercept=False)),
    "edge_knots": dict(model="LinearGAM", terms=[s(0, edge_knots=[0.1, 0.8]), s(1)],
                       data=("c4", 1200)),
    "logistic_few_iter": dict(model="LogisticGAM", terms=[s(0), s(1)], data=("c2", 1500),
                              model_kwargs=dict(max_iter=2)),
    "gam_generic": dict(model="GAM", terms=[s(0), s(1)], data=("positive", 1200),
                        model_kwargs=dict(distribution="gamma", link="inverse")),
    "lam_large": dict(model="LogisticGAM", terms=[s(0, lam=1e3), s(1, lam=1e-3)],
                      data=("c2", 1500)),
}


GRID_CASES = {
    # default 11-point grid on config 1 (docs/notebooks/tour_of_pygam.ipynb:375-380)
    "c1_wage_default_grid": dict(fit="c1_wage", grid=None),
    # config 4 shape; 5^3 cartesian grid keeps the fixture cheap, 11^3 is exercised in the bench
    "c4_grid5": dict(fit="c4_linear", grid=dict(lam=[list(np.logspace(-3, 3, 5))] * 3)),
    "c4_grid_rows": dict(fit="c4_linear",
                         grid=dict(lam=np.exp(np.random.default_rng(3).random((6, 3)) * 6 - 3))),
    "logistic_grid": dict(fit="lam_large", grid=None),
}


# b_spline_basis known-answer vectors: (n_splines, spline_order, periodic, edge_knots)
BASIS_CASES = {
    "cubic20": (20, 3, False, (0.0, 1.0)),
    "cubic25_shift": (25, 3, False, (-3.0, 7.5)),
    "cubic4_min": (4, 3, False, (0.0, 1.0)),
    "quad12": (12, 2, False, (0.0, 2.0)),
    "lin9": (9, 1, False, (0.0, 1.0)),
    "haar6": (6, 0, False, (-0.5, 5.5)),
    "quartic7": (7, 4, False, (0.0, 1.0)),
    "cyc10": (10, 3, True, (0.0, 1.0)),
    "cyc8_o2": (8, 2, True, (0.0, 6.0)),
    "degenerate": (6, 3, False, (2.0, 2.0)),
}


def basis_x(edge_knots, periodic=False):
    """query points: interior grid, exact knots/edges, and points outside the edge knots.

    For the cyclic basis the reference raises a shape error for normalised x in (1, 1+1e-9]
    (``utils.py:694`` wraps modulo 1+1e-9, then ``utils.py:762`` extrapolates with the un-folded
    gradient table), so that sliver is left out of the periodic query set.
    """
    a, b = edge_knots
    w = (b - a) if b > a else 1.0
    grid = np.linspace(a, a + w, 97)
    extra = [a - 0.35 * w, a, a + 1e-15, a + w, a + 1.2 * w, a + 0.5 * w,
             a + w / 3.0, a + 2.75 * w, a - 1.5 * w]
    if not periodic:
        extra += [a + w * (1 + 1e-12), a - 1e-12 * w]
    return np.concatenate([grid, np.array(extra)])


def make_data(tag, wage_xy=None):
    """returns dict(X, y[, exposure])"""
    kind = tag[0]
    if kind == "wage":
        X, y = wage_xy
        return dict(X=np.asarray(X, dtype=np.float64), y=np.asarray(y, dtype=np.float64))
    n = tag[1]
    if kind == "c2":
        X, y = data_c2(n)
    elif kind == "c3":
        X, y = data_c3(n)
    elif kind == "c4":
        X, y = data_c4(n)
    elif kind == "c5":
        X, y = data_c5(n)
    elif kind == "positive":
        X, y = data_positive(n)
    elif kind == "mixed":
        X, y = data_mixed(n)
    elif kind == "counts_exposure":
        X, y, e = data_counts_exposure(n)
        return dict(X=X, y=y, exposure=e)
    elif kind == "counts_exposure2":
        X, y, _ = data_counts_exposure(n, seed=9)
        return dict(X=X, y=y)
    else:
        raise KeyError(kind)
    return dict(X=X, y=y)


def make_weights(kind, n):
    if kind is None:
        return None
    rng = np.random.default_rng(99)
    if kind == "uniform":
        return rng.random(n) * 2.0 + 0.25
    raise KeyError(kind)
