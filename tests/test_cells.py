"""The cell-owned Gram pass (pgb_cells.cu) of all-spline models: its host plan (CPU) and, on the
GPU, agreement with the general shared-memory path of pgb_gram.cu and with the oracle on the same
inputs, over the shapes that exercise its scheduling (odd n, several tiles and bands, role
switches, key ranges, extrapolated rows, weights, no intercept)."""
import contextlib
import io

import numpy as np
import pytest

import cases
import helpers
import pygam_b200 as pg
from pygam_b200 import _lib
from test_plan import plan_for


def test_cell_plan_of_all_spline_models():
    rng = np.random.default_rng(0)
    _, info, errs = plan_for(cases.FIT_CASES["c2_logistic"], rng.random((50, 10)))
    assert errs == 0
    assert info.cell_roles == 55 and info.cell_tile_rows >= 1024  # 45 pairs + 10 diagonal blocks
    # 22 x 22 knot cells per pair: 512-thread CTAs, the ring kernel, no generic roles
    assert (info.cell_threads, info.cell_ring_groups, info.cell_generic_roles) == (512, 1, 0)
    # f(), l() and te() of two cubic marginals are generic pairs of the same pass (configs 1 and 3) ...
    _, info1, e1 = plan_for(cases.FIT_CASES["c1_wage"], helpers.load("fit_c1_wage")["X"])
    _, info3, e3 = plan_for(cases.FIT_CASES["c3_poisson"], rng.random((50, 24)))
    assert info1.cell_roles > 3 and info3.cell_roles > 210 + 40 and (e1, e3) == (0, 0)
    # config 3: 17 x 17 cells per spline pair (320-thread CTAs, two per SM), te() pairs in the generic kernel
    assert (info3.cell_threads, info3.cell_ring_groups) == (320, 0) and info3.cell_generic_roles > 40
    # ... other spline orders, cyclic bases and by= keep the general path
    _, infom, em = plan_for(cases.FIT_CASES["mixed_terms"], helpers.load("fit_mixed_terms")["X"])
    assert (infom.cell_roles, em) == (0, 0)


@pytest.mark.parametrize("n_splines", [4, 5, 20, 60])
def test_cell_plan_covers_every_entry(n_splines):
    """key ranges: a pair with more than 512 cells is served by several roles"""
    case = {"model": "LinearGAM", "terms": [dict(kind="s", feature=0, n_splines=n_splines),
                                            dict(kind="s", feature=1, n_splines=max(4, n_splines // 2))]}
    rng = np.random.default_rng(1)
    _, info, errs = plan_for(case, rng.random((40, 2)))
    assert errs == 0 and info.cell_roles >= 3


# ---------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


def _engines(gam, n_features, env):
    """two engines of the same compiled model: cell-owned Gram pass and the general path (the test-only entry
    point pgb_debug_model_create constrains the plan)"""
    from pygam_b200.engine import Engine

    def make(cells):
        opts = _lib.PgbDebugOpts(0 if cells else 1, int(env.get("tile", 0)), int(env.get("band", 0)), int(env.get("threads", 0)), int(env.get("kernel", 0)))
        return Engine(gam.terms, gam.distribution.name, gam.link.name, gam.distribution.levels, gam._expectile, n_features,
                      debug_opts=opts)

    a, b = make(True), make(False)
    assert a.info.cell_roles > 0 and b.info.cell_roles == 0
    return a, b


def _compare(gam, X, y, w=None, env=None, modes=(_lib.MODE_IRLS, _lib.MODE_INIT)):
    import torch
    from pygam_b200.engine import DeviceData

    cells, general = _engines(gam, X.shape[1], env or {})
    # a model with generic pairs runs the same prepass with and without the stored order: bit-equal results
    mixed = any(type(t).__name__ in ("TensorTerm", "FactorTerm", "LinearTerm") for t in gam.terms)
    data = DeviceData.from_host(X, y, None if w is None else np.asarray(w, dtype=np.float32))
    m = cells.m
    rng = np.random.default_rng(3)
    coef = np.asarray(gam.coef_) * (1 + 0.05 * rng.standard_normal(m))
    for mode in modes:
        a = cells.gram_pass(data, coef, mode=mode).cpu().numpy().copy()
        b = general.gram_pass(data, coef, mode=mode).cpu().numpy().copy()
        G, Gr = a[:m * m].reshape(m, m), b[:m * m].reshape(m, m)
        scale = np.abs(Gr).max()
        assert np.abs(G - Gr).max() <= 2e-13 * scale, (mode, np.abs(G - Gr).max() / scale)
        np.testing.assert_array_equal(G, G.T)
        assert np.abs(a[m * m:m * m + m] - b[m * m:m * m + m]).max() <= 2e-13 * max(scale, np.abs(b[m * m:m * m + m]).max())
        sa, sb = a[m * m + m:], b[m * m + m:]
        for s in (_lib.GS_NUSED, _lib.GS_NROWS, _lib.GS_NCORRECT):
            assert sa[s] == sb[s]
        assert abs(sa[_lib.GS_DEV] - sb[_lib.GS_DEV]) <= 1e-12 * max(1.0, abs(sb[_lib.GS_DEV]))
        again = cells.gram_pass(data, coef, mode=mode).cpu().numpy()
        assert np.array_equal(a, again), "the cell-owned pass must be bit-reproducible"
        # the prepared order (pgb_gram_prepare, the default for resident data) and the in-kernel sort of the
        # streaming path visit the rows of a cell in the same order; the prepared prepass rebuilds t from the
        # stored centred coordinate (one rounding in lp), so IRLS weights agree to rounding, INIT bit for bit
        cells.use_plan = False
        try:
            unplanned = cells.gram_pass(data, coef, mode=mode).cpu().numpy()
        finally:
            cells.use_plan = True
        # same sums, same order of the rows inside a cell; the prepared prepass of an all-spline model rebuilds t from
        # the stored centred coordinate (one rounding in lp): agreement to rounding
        assert np.abs(a - unplanned).max() <= 1e-13 * scale
        # the optional second step of the preparation (pgb_gram_refine) reorders the rows inside every cell
        cells.refine_order(data)
        refined = cells.gram_pass(data, coef, mode=mode).cpu().numpy().copy()
        assert np.abs(refined - b).max() <= 2e-13 * max(scale, np.abs(b[m * m:m * m + m]).max())
        assert np.array_equal(refined, cells.gram_pass(data, coef, mode=mode).cpu().numpy())
        cells._plan_key = None  # the next mode starts from a fresh preparation
    del torch


def _fit_small(case, X, y, **kw):
    gam = helpers.build_model(case, max_iter=3, **kw)
    with contextlib.redirect_stdout(io.StringIO()):
        gam.fit(X[:4000], y[:4000])
    return gam


@gpu
@pytest.mark.parametrize("n,env", [(20001, {}), (30000, {"tile": 1024, "band": 2}), (30000, {"tile": 1024, "band": 2, "kernel": 1}),
                                   (50001, {"tile": 512, "band": 3, "kernel": 2, "threads": 320}),
                                   (777, {}), (65536, {"tile": 2048, "band": 1})])
def test_cells_equal_general_path_config2(n, env):
    X, y = cases.data_c2(n, seed=11)
    X[:4000] = np.clip(X[:4000], 0.0, 1.0)
    X[0, :] = X.min(axis=0)  # the compiled edge knots (from the first 4000 rows) span all rows
    X[1, :] = X.max(axis=0)
    gam = _fit_small(cases.FIT_CASES["c2_logistic"], X, y)
    _compare(gam, X, y, env=env)


@gpu
def test_cells_extrapolated_rows_weights_and_skewed_columns():
    """rows outside the edge knots take eval_marg's linear extrapolation; normal columns make the
    cell counts very uneven"""
    rng = np.random.default_rng(5)
    n = 25_001
    X = rng.standard_normal((n, 3))
    X[4000:] *= 1.7  # later rows reach beyond the knots compiled from the first 4000
    eta = np.sin(X[:, 0]) + 0.3 * X[:, 1] ** 2 - 0.5 * np.cos(2 * X[:, 2])
    y = rng.poisson(np.exp(0.3 * eta)).astype(float)
    w = rng.uniform(0.5, 2.0, n)
    case = {"model": "PoissonGAM", "terms": [dict(kind="s", feature=0, n_splines=12), dict(kind="s", feature=1),
                                             dict(kind="s", feature=2, n_splines=31)]}
    gam = _fit_small(case, X, y)
    _compare(gam, X, y, w=w, env={"tile": 4096})


@gpu
def test_cells_key_ranges_and_no_intercept():
    rng = np.random.default_rng(6)
    n = 40_000
    X = rng.random((n, 2))
    X[0], X[1] = 0.0, 1.0
    y = np.sin(6 * X[:, 0]) + X[:, 1] + 0.1 * rng.standard_normal(n)
    case = {"model": "LinearGAM", "terms": [dict(kind="s", feature=0, n_splines=60), dict(kind="s", feature=1, n_splines=33)]}
    gam = _fit_small(case, X, y, fit_intercept=False)
    assert gam.terms.n_coefs == 93
    _compare(gam, X, y, modes=(_lib.MODE_IRLS,))
    gam1 = _fit_small({"model": "GammaGAM", "terms": [dict(kind="s", feature=0, n_splines=9)]}, X, np.exp(y))
    _compare(gam1, X, np.exp(y))


@gpu
@pytest.mark.parametrize("name,env", [("c1_wage", {}), ("c1_wage", {"tile": 512, "band": 2}), ("c3_poisson", {}),
                                      ("c3_poisson", {"tile": 1024, "band": 1, "threads": 256}), ("dummy_coding", {}),
                                      ("multi_penalty", {"tile": 256})])
def test_cells_equal_general_path_mixed_terms(name, env):
    """te() / f() / l() blocks as generic cell pairs (BASELINE configs 1 and 3) vs the general path"""
    case = cases.FIT_CASES[name]
    g = helpers.load("fit_" + name)
    gam = helpers.build_model(case, max_iter=3)
    with contextlib.redirect_stdout(io.StringIO()):
        if case["model"] == "PoissonGAM":
            gam.fit(g["X"], g["y"], exposure=g.get("exposure"), weights=g.get("weights"))
        else:
            gam.fit(g["X"], g["y"], weights=g.get("weights"))
    _compare(gam, g["X"], g["y"], w=g.get("weights"), env=env)


@gpu
def test_cells_te_f_l_rows_outside_knots_large():
    """te() twice (4-d cells), f(), l(), custom edge knots narrower than the data, several tiles and bands"""
    rng = np.random.default_rng(7)
    n = 50_003
    X = np.column_stack([rng.random(n), rng.random(n), rng.integers(0, 4, n).astype(float), rng.standard_normal(n),
                         rng.random(n), rng.random(n)])
    eta = np.sin(4 * X[:, 0]) * X[:, 1] + 0.3 * X[:, 2] + 0.2 * X[:, 3] + X[:, 4] * X[:, 5]
    y = rng.poisson(np.exp(0.4 * eta)).astype(float)
    s, te, f, l = cases.s, cases.te, cases.f, cases.l
    case = dict(model="PoissonGAM",
                terms=[te(s(0, n_splines=6), s(1, n_splines=5)), f(2), l(3), s(4, n_splines=7, edge_knots=[0.1, 0.8]),
                       te(s(5, n_splines=5, edge_knots=[0.2, 0.9]), s(0, n_splines=4))])
    gam = _fit_small(case, X, y)
    _compare(gam, X, y, env={"tile": 2048, "band": 3})


@gpu
@pytest.mark.parametrize("env", [{}, {"tile": 1024, "band": 2}])
def test_cells_by_variables_and_linear_marginals(env):
    """s(j, by=k), te(s, l) and te(l, s) are generic pairs of the cell-owned pass (a cubic factor times a value factor):
    equal to the general path, with weights, rows outside the knots, several tiles"""
    rng = np.random.default_rng(11)
    n = 40_001
    # every by-variable / l() column is its own feature: the rows of s(j, by=k) and of te(s, l(k)) sum to x_k, so a column
    # shared between two such terms (or with an l()) would make the design exactly collinear
    X = np.column_stack([rng.random(n), 2.0 * rng.random(n) - 0.5, rng.integers(0, 3, n).astype(float), rng.standard_normal(n),
                         rng.random(n), rng.random(n), rng.standard_normal(n), rng.random(n) + 0.5, rng.standard_normal(n)])
    eta = np.sin(4 * X[:, 0]) * X[:, 1] + 0.3 * X[:, 2] + 0.2 * X[:, 3] * X[:, 4] + X[:, 5] + 0.1 * X[:, 6] * X[:, 5] + 0.2 * X[:, 8]
    y = rng.poisson(np.exp(0.3 * eta)).astype(float)
    w = rng.uniform(0.5, 2.0, n)
    s, te, f, l = cases.s, cases.te, cases.f, cases.l
    case = dict(model="PoissonGAM",
                terms=[s(0, by=1, n_splines=7), te(s(4, n_splines=6), l(3)), te(l(6), s(5, n_splines=5, edge_knots=[0.1, 0.85])),
                       s(5, n_splines=8), f(2), l(8), te(s(0, n_splines=5), s(4, n_splines=4)), s(4, by=7, n_splines=6)])
    gam = _fit_small(case, X, y)
    _compare(gam, X, y, w=w, env=env)
