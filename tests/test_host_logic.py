"""CPU tests of the product's host logic (no GPU): the PIRLS driver, penalties / constraints,
statistics assembly and grid search of pygam_b200 are run with the oracle-backed test engine
(tests/oracle_backend.py) and compared with the golden fixtures of the unmodified reference.
The solve used here is a numpy mirror of the CUDA solver's algorithm, so these tests also show
that the Gram-form, null-space-aware method reproduces the reference's iteration counts."""
import numpy as np
import pytest

import cases
import helpers
from helpers import check_against_golden, fit_case, load
from oracle_backend import OracleBackend
from pygam_b200 import engine as pg_engine
from pygam_b200 import penalties as pg_pen
import pygam_b200 as pg


@pytest.fixture(autouse=True)
def oracle_backend():
    old = pg_engine.swap_backend_for_tests(OracleBackend())
    yield
    pg_engine.swap_backend_for_tests(old)


def test_penalties_known_answers():
    g = load("penalties")
    for n in (1, 2, 3, 5, 8, 20):
        coef = g[f"coef_{n}"]
        np.testing.assert_array_equal(pg_pen.derivative(n).toarray(), g[f"derivative_{n}"])
        np.testing.assert_array_equal(pg_pen.l2(n).toarray(), g[f"l2_{n}"])
        if n >= 5:
            np.testing.assert_array_equal(pg_pen.periodic(n).toarray(), g[f"periodic_{n}"])
        for cname in ("monotonic_inc", "monotonic_dec", "convex", "concave"):
            np.testing.assert_array_equal(pg_pen.CONSTRAINTS[cname](n, coef).toarray(), g[f"{cname}_{n}"])


@pytest.mark.parametrize("name", sorted(cases.FIT_CASES))
def test_fit_matches_reference(name):
    case = cases.FIT_CASES[name]
    g = load("fit_" + name)
    gam = fit_case(case, g)
    np.testing.assert_array_equal(gam._P(), g["penalty"])
    check_against_golden(gam, g, constrained=name in helpers.CONSTRAINED)
    if "constraint_final" in g:
        C = gam.terms._constraints(g["coef"], gam._constraint_lam, gam._constraint_l2)
        np.testing.assert_allclose(C, g["constraint_final"], rtol=1e-15, atol=0)


@pytest.mark.parametrize("name", sorted(cases.GRID_CASES))
def test_gridsearch_matches_reference(name):
    gcase = cases.GRID_CASES[name]
    case = cases.FIT_CASES[gcase["fit"]]
    g = load("grid_" + name)
    f = load("fit_" + gcase["fit"])
    gam = helpers.build_model(case)
    kw = {} if gcase["grid"] is None else dict(gcase["grid"])
    if case["model"] == "PoissonGAM":
        scores = gam.gridsearch(f["X"], f["y"], exposure=f.get("exposure"), weights=f.get("weights"),
                                return_scores=True, **kw)
    else:
        scores = gam.gridsearch(f["X"], f["y"], weights=f.get("weights"), return_scores=True, **kw)
    models = list(scores.keys())
    vals = np.array(list(scores.values()))
    assert len(vals) == len(g["scores"])
    np.testing.assert_allclose(vals, g["scores"], rtol=1e-9)
    np.testing.assert_array_equal([len(mm.logs_["diffs"]) for mm in models], g["n_iters"])
    np.testing.assert_allclose([mm.statistics_["edof"] for mm in models], g["edofs"], rtol=1e-8)
    obj = "UBRE" if gam.distribution._known_scale else "GCV"
    assert abs(gam.statistics_[obj] - float(g["best_score"])) <= 1e-10 * abs(float(g["best_score"]))


def test_config1_published_known_answers():
    """docs/notebooks/quick_start.ipynb:113-118 and tour_of_pygam.ipynb:375-380"""
    w = load("wage_xy")
    gam = pg.LinearGAM(pg.s(0) + pg.s(1) + pg.f(2)).fit(w["X"], w["y"])
    st = gam.statistics_
    assert round(st["edof"], 4) == 25.1911
    assert round(st["GCV"], 4) == 1255.6902
    assert round(st["pseudo_r2"]["explained_deviance"], 4) == 0.2955
    assert len(gam.logs_["diffs"]) == 2
    gam2 = pg.LinearGAM(pg.s(0) + pg.s(1) + pg.f(2)).gridsearch(w["X"], w["y"])
    assert round(gam2.lam[0][0], 4) == 15.8489
    assert round(gam2.statistics_["edof"], 4) == 19.2602
    assert round(gam2.statistics_["GCV"], 4) == 1250.3656


def test_term_grammar_and_validation():
    tl = pg.s(0) + pg.s(1, n_splines=7) + pg.f(2) + pg.te(3, 4) + pg.l(5)
    assert len(tl) == 5
    assert repr(tl) == "s(0) + s(1) + f(2) + te(3, 4) + l(5)"
    assert len(pg.s(0) + pg.s(0)) == 1  # duplicates are dropped (terms.py:1646)
    with pytest.raises(ValueError):
        pg.s(0, n_splines=3, spline_order=3)
    with pytest.raises(ValueError):
        pg.s(0, basis="xx")
    with pytest.raises(ValueError):
        pg.te(0)
    with pytest.raises(ValueError):
        pg.s(0, lam=-1)
    with pytest.raises(ValueError):
        pg.f(0, coding="weird")
    assert pg.te(0, 1).n_coefs == 100
    assert pg.s(0, constraints="none").hasconstraint  # the string counts (terms.py:260-263)


def test_input_validation():
    rng = np.random.default_rng(0)
    X = rng.random((50, 2))
    y = rng.random(50)
    Xn = X.copy()
    Xn[3, 1] = np.nan
    with pytest.raises(ValueError):
        pg.LinearGAM().fit(Xn, y)
    yn = y.copy()
    yn[0] = np.inf
    with pytest.raises(ValueError):
        pg.LinearGAM().fit(X, yn)
    with pytest.raises(ValueError):
        pg.LogisticGAM().fit(X, y + 2.0)  # outside the logit domain
    with pytest.raises(ValueError):
        pg.LinearGAM().fit(X, y[:-1])
    with pytest.raises(ValueError):
        pg.LinearGAM(pg.s(5)).fit(X, y)
    with pytest.raises(NotImplementedError):
        pg.LinearGAM(callbacks=[object()])
    with pytest.raises(AttributeError):
        pg.LinearGAM().predict_mu(X)


@pytest.mark.parametrize("name", ["c1_wage", "c3_poisson", "c5_expectile", "concave_dec", "multi_penalty", "tensor_by"])
def test_public_sparse_penalties_equal_the_dense_ones_of_the_engine(name):
    """build_penalties / build_constraints (scipy sparse, the reference's return type, terms.py:307-396, 1479-1609,
    1925-1979) against the dense matrices the engine uses; TensorTerm's marginal helpers sum to the term's matrix"""
    import scipy.sparse

    g = load("fit_" + name)
    gam = helpers.build_model(cases.FIT_CASES[name])
    gam.terms.compile(g["X"])
    coef = np.sin(np.arange(gam.terms.n_coefs) * 1.7)
    P = gam.terms.build_penalties()
    C = gam.terms.build_constraints(coef, 1e9, 1e-3)
    assert scipy.sparse.issparse(P) and scipy.sparse.issparse(C)
    np.testing.assert_array_equal(P.toarray(), gam.terms._penalties())
    np.testing.assert_array_equal(C.toarray(), gam.terms._constraints(coef, 1e9, 1e-3))
    for i, t in enumerate(gam.terms):
        sl = gam.terms.get_coef_indices(i)
        np.testing.assert_array_equal(t.build_penalties().toarray(), t._penalties())
        if t.kind == "tensor_term":
            tot = sum(t._build_marginal_penalties(q).toarray() for q in range(len(t._terms)))
            np.testing.assert_array_equal(tot, t._penalties())
            totc = sum(t._build_marginal_constraints(q, coef[sl], 1e9, 1e-3).toarray() for q in range(len(t._terms)))
            np.testing.assert_array_equal(totc, t._constraints(coef[sl], 1e9, 1e-3))


def test_model_repr_follows_the_reference_format():
    """Core.__repr__ / nice_repr (core.py:7-138): class name, the user-facing parameters in alphabetical order, floats
    rounded to four decimals, wrapped at 70 columns with a three-space indent (strings recorded from pyGAM 0.12.0)"""
    from pygam_b200.utils import nice_repr

    g = pg.LinearGAM(pg.s(0, n_splines=7) + pg.te(1, 2) + pg.f(3), lam=0.5)
    assert repr(g) == ("LinearGAM(callbacks=['deviance', 'diffs'], fit_intercept=True, \n"
                       "   lam=0.5, max_iter=100, scale=None, terms=s(0) + te(1, 2) + f(3), \n"
                       "   tol=0.0001, verbose=False)")
    assert repr(pg.LogisticGAM()) == ("LogisticGAM(callbacks=['deviance', 'diffs', 'accuracy'], \n"
                                      "   fit_intercept=True, max_iter=100, terms='auto', tol=0.0001, \n"
                                      "   verbose=False)")
    assert nice_repr("hi", {}, line_width=30, line_offset=5, decimals=3) == "hi()"
    assert nice_repr("hi", {"color": "blue", "n_ears": 3, "height": 1.3336}, line_width=60, line_offset=5,
                     decimals=3) == "hi(color='blue', height=1.334, n_ears=3)"


def test_distribution_and_link_objects_by_the_reference_names():
    """pygam_b200.distributions / pygam_b200.links (distributions.py:97-621, links.py:21-323): the named classes
    configure a GAM like the strings do and carry the reference's host-side methods (closed forms checked here; all of
    them were compared with the reference's classes on random inputs when they were written)"""
    import scipy.stats

    from pygam_b200.distributions import DISTRIBUTIONS, BinomialDist, GammaDist, NormalDist, PoissonDist
    from pygam_b200.links import LINKS, LogitLink, LogLink

    assert sorted(DISTRIBUTIONS) == ["binomial", "gamma", "inv_gauss", "normal", "poisson"]
    assert sorted(LINKS) == ["identity", "inv_squared", "inverse", "log", "logit"]
    mu, y, w = np.array([0.5, 1.5, 2.0]), np.array([1.0, 1.0, 3.0]), np.array([1.0, 2.0, 0.5])
    d = BinomialDist(levels=3)
    np.testing.assert_allclose(d.V(mu), mu * (1 - mu / 3))
    np.testing.assert_allclose(d.V(mu, weights=w), mu * (1 - mu / 3) / w)
    np.testing.assert_allclose(d.log_pdf(y, mu), scipy.stats.binom.logpmf(y, 3, mu / 3))
    np.testing.assert_allclose(LogitLink().mu(LogitLink().link(mu, d), d), mu)
    np.testing.assert_allclose(LogitLink().gradient(mu, d), 3 / (mu * (3 - mu)))
    n = NormalDist(scale=2.0)
    np.testing.assert_allclose(n.deviance(y, mu), (y - mu) ** 2 / 4.0)
    np.testing.assert_allclose(n.deviance(y, mu, scaled=False, weights=w), (y - mu) ** 2 * w)
    assert n.phi(y, mu, 1.0, w) == 4.0 and not GammaDist()._known_scale
    np.testing.assert_allclose(PoissonDist().log_pdf(y, mu, weights=w), scipy.stats.poisson.logpmf(y, mu * w))
    g1 = pg.GAM(pg.s(0), distribution=GammaDist(), link=LogLink())
    g2 = pg.GAM(pg.s(0), distribution="gamma", link="log")
    assert (g1.distribution.name, g1.link.name) == (g2.distribution.name, g2.link.name) == ("gamma", "log")
    assert pg.LogisticGAM().distribution.V(np.array([0.25]))[0] == 0.1875


def test_callback_objects_by_the_reference_names():
    """pygam_b200.callbacks (callbacks.py:98-236): the four built-in classes are accepted like their names; any other
    CallBack is refused (it would need the n-sized y / mu of every iteration on the host)"""
    from pygam_b200.callbacks import CALLBACKS, Accuracy, CallBack, Coef, Deviance, Diffs

    assert sorted(CALLBACKS) == ["accuracy", "coef", "deviance", "diffs"] and str(Deviance()) == "deviance"
    g = pg.LogisticGAM(pg.s(0), callbacks=[Deviance(), Diffs(), Accuracy(), Coef()])
    g._validate_params()
    assert g.callbacks == ["deviance", "diffs", "accuracy", "coef"]

    class Mine(CallBack):
        def on_loop_start(self, y, mu):
            return 0

    with pytest.raises(NotImplementedError):
        pg.LinearGAM(pg.s(0), callbacks=[Mine(name="mine")])._validate_params()
