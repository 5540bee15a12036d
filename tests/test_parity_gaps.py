"""Parity cases round 1 left open (VERDICT r1 "What's weak" 1-3): masked rows (GAM._mask, pygam.py:638-663 -- logit
overflow to NaN, weights below sqrt(EPS)), the all-masked OptimizationError, a non-positive-definite penalty skipped
inside gridsearch (pygam/tests/test_utils.py:176-183), the inv_squared link, p-values / standard errors / AICc,
ExpectileGAM.fit_quantile, pickle / deepcopy of a fitted model, rows < coefficients, the 11^3 lam grid of config 4 and
configs 2 / 3 at 1e5 / 5e4 rows.  Fixtures: tests/golden/r2_*.npz, outputs of the unmodified reference
(tests/golden/make_golden_r2.py).  Every product test runs twice: on the CPU with the oracle-backed test engine (host
logic) and, marked gpu, on the CUDA engine through the C ABI."""
import contextlib
import copy
import hashlib
import io
import os
import pickle
import warnings

import numpy as np
import pytest

import cases
import helpers
from helpers import rel
from oracle import gam_oracle as orc
from oracle_backend import OracleBackend
import pygam_b200 as pg
from pygam_b200 import engine as pg_engine

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BACKENDS = [pytest.param("oracle", id="host-logic"), pytest.param("cuda", marks=pytest.mark.gpu, id="cuda")]


def r2(name):
    return dict(np.load(os.path.join(GOLD, f"r2_{name}.npz")))


@contextlib.contextmanager
def backend(kind):
    if kind == "oracle":
        old = pg_engine.swap_backend_for_tests(OracleBackend())
        try:
            yield
        finally:
            pg_engine.swap_backend_for_tests(old)
    else:
        yield


def quiet_fit(gam, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return gam.fit(*a, **kw)


def check_fit(gam, g, X, coef_tol=1e-8, stat_tol=1e-10, log_tol=1e-8):
    assert len(gam.logs_["diffs"]) == int(g["n_iter"])
    assert (gam.logs_["diffs"][-1] < gam.tol) == bool(g["log_diffs"][-1] < g["tol"])
    np.testing.assert_allclose(gam.logs_["deviance"], g["log_deviance"], rtol=log_tol)
    Pz = helpers.identifiable_projector(gam)
    err = np.linalg.norm(Pz @ (gam.coef_ - g["coef"])) / np.linalg.norm(g["coef"])
    assert err < coef_tol, err
    st = gam.statistics_
    for key in ("edof", "deviance", "scale", "loglikelihood", "AIC", "AICc"):
        refv = float(g["stat_" + key])
        assert abs(st[key] - refv) <= max(stat_tol, 1e-9 if key in ("edof", "loglikelihood", "AIC", "AICc") else 0) * max(1.0, abs(refv)), (key, st[key], refv)
    for key in ("GCV", "UBRE"):
        refv = float(g["stat_" + key])
        if not np.isnan(refv):
            assert abs(st[key] - refv) <= stat_tol * abs(refv), (key, st[key], refv)
    assert rel(gam._linear_predictor(X[:64]), g["lp_head"]) < 1e-8


# ---------------------------------------------------------------------------------------------
# the oracle itself against the new fixtures
# ---------------------------------------------------------------------------------------------
MASK_LOGIT = dict(model="LogisticGAM", terms=[cases.s(0), cases.s(1), cases.s(2)])
INV_SQ = dict(model="GAM", terms=[cases.s(0), cases.s(1), cases.s(2)], model_kwargs=dict(distribution="gamma", link="inv_squared"))


def _oracle_check(res, g, tol=1e-9):
    # the reference creates logs_ once (pygam.py:901-902): a refitted object carries the logs of its earlier fits in
    # front of the last fit's
    k = len(res["logs"]["diffs"])
    assert k <= int(g["n_iter"])
    np.testing.assert_allclose(res["logs"]["deviance"], g["log_deviance"][-k:], rtol=1e-9)
    for key in ("edof", "deviance", "AIC", "AICc", "loglikelihood"):
        assert abs(res["statistics"][key] - float(g["stat_" + key])) <= tol * max(1.0, abs(float(g["stat_" + key]))), key


def test_oracle_masked_rows_logit_overflow():
    g = r2("mask_logit")
    with np.errstate(all="ignore"):
        res = orc.fit(MASK_LOGIT, g["X"], g["y"], coef=g["coef0"])
    _oracle_check(res, g)
    # half of the rows overflow in the first iteration: exp(lp) = inf, mu = NaN (links.py:121-122)
    lp0 = orc.model_matrix(orc.compile_terms(orc.normalize_terms(MASK_LOGIT["terms"]), g["X"]), g["X"]).dot(g["coef0"])
    assert (lp0 > 709.8).sum() > 1000


def test_oracle_masked_rows_tiny_weights_and_inv_squared_and_truncated_svd():
    g = r2("mask_weights")
    _oracle_check(orc.fit(cases.FIT_CASES["c4_linear"], g["X"], g["y"], weights=g["weights"]), g)
    g = r2("inv_squared")
    with np.errstate(all="ignore"):
        _oracle_check(orc.fit(INV_SQ, g["X"], g["y"]), g, tol=1e-7)
    g = r2("rows_lt_coefs")  # n = 15 < m = 41: the truncated SVD of pygam.py:789-799 is part of the restatement
    res = orc.fit(dict(model="LinearGAM", terms=[cases.s(0), cases.s(1)]), g["X"], g["y"])
    assert len(res["logs"]["diffs"]) == int(g["n_iter"])
    assert abs(res["statistics"]["edof"] - float(g["stat_edof"])) < 1e-6


TE_LINEAR = dict(model="PoissonGAM", terms=[cases.te(cases.s(0, n_splines=8), cases.l(1), lam=[0.3, 2.0]), cases.s(2, n_splines=12),
                                              cases.te(cases.l(3), cases.s(2, n_splines=6))])


def test_oracle_tensor_terms_with_linear_marginals():
    """te(s, l) and te(l, s) (terms.py:1330-1356: a marginal is any non-tensor term; pygam/tests/test_terms.py:183-199)"""
    g = r2("te_linear_marginal")
    terms = orc.compile_terms(orc.normalize_terms(TE_LINEAR["terms"]), g["X"])
    np.testing.assert_array_equal(orc.penalty_matrix(terms), g["penalty"])
    assert rel(orc.model_matrix(terms, g["X"][:32]).toarray(), g["modelmat_head"]) < 1e-14
    _oracle_check(orc.fit(TE_LINEAR, g["X"], g["y"]), g)


# ---------------------------------------------------------------------------------------------
# the product
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", BACKENDS)
def test_masked_rows_logit_overflow(kind):
    """lp > 709 on half of the rows of the first iteration: mu = NaN, the rows are masked out of G, r, the
    logged deviance and the accuracy; the fit then recovers like the reference's (12 iterations)"""
    g = r2("mask_logit")
    with backend(kind):
        gam = helpers.build_model(MASK_LOGIT)
        quiet_fit(gam, g["X"], g["y"])
        gam.coef_ = g["coef0"].copy()
        quiet_fit(gam, g["X"], g["y"])
        # iterations 6-9 solve a rank-57 system of 57 kept rows whose cropped spectrum reaches down to the
        # sqrt(EPS)-sized eigenvalues: their logged deviances agree to 1e-7 only, in the reference's own arithmetic
        # too; the converged fit is held to the usual tolerances
        if kind == "oracle":
            check_fit(gam, g, g["X"], log_tol=1e-6)
            np.testing.assert_allclose(gam.logs_["accuracy"], g["log_accuracy"], rtol=0, atol=1e-12)
            return
        # On the device the kept set itself is not reproducible: rows with lp near 35 have W = (g'^2 V)^-1/2 within
        # an ulp of exp() of the sqrt(EPS) threshold of GAM._mask, CUDA's exp and glibc's differ in the last bit,
        # and one row more or less changes the RANK of the cropped solve.  What must hold: the same masked start
        # (logged deviance of the kept rows of iteration 6), recovery, and the reference's converged fit.
        assert abs(len(gam.logs_["diffs"]) - int(g["n_iter"])) <= 2 and gam.logs_["diffs"][-1] < gam.tol
        np.testing.assert_allclose(gam.logs_["deviance"][:6], g["log_deviance"][:6], rtol=1e-9)
        st = gam.statistics_
        assert abs(st["UBRE"] - float(g["stat_UBRE"])) < 1e-7 * float(g["stat_UBRE"])
        assert abs(st["edof"] - float(g["stat_edof"])) < 1e-4
        assert rel(gam._linear_predictor(g["X"][:64]), g["lp_head"]) < 1e-4


@pytest.mark.parametrize("kind", BACKENDS)
def test_masked_rows_tiny_weights_and_all_masked(kind):
    g = r2("mask_weights")
    with backend(kind):
        gam = helpers.build_model(cases.FIT_CASES["c4_linear"])
        quiet_fit(gam, g["X"], g["y"], weights=g["weights"])
        check_fit(gam, g, g["X"])
        # every row below the mask threshold: OptimizationError (pygam.py:658-662)
        with pytest.raises(pg.OptimizationError):
            quiet_fit(helpers.build_model(cases.FIT_CASES["c4_linear"]), g["X"], g["y"], weights=np.full(len(g["y"]), 1e-20))


@pytest.mark.parametrize("kind", BACKENDS)
def test_inv_squared_link(kind):
    g = r2("inv_squared")
    with backend(kind):
        gam = pg.GAM(pg.s(0) + pg.s(1) + pg.s(2), distribution="gamma", link="inv_squared")
        quiet_fit(gam, g["X"], g["y"])
        check_fit(gam, g, g["X"], coef_tol=1e-7, stat_tol=1e-8)


@pytest.mark.parametrize("kind", BACKENDS)
def test_gridsearch_skips_a_penalty_that_is_not_positive_definite(kind):
    """pygam/tests/test_utils.py:176-183: lam = 1e10 .. 1e12 must not crash the grid search; a candidate whose
    S + P has no Cholesky factor is skipped (pygam.py:2055-2066)"""
    X, y = cases.data_c2(1500)
    with backend(kind), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gam = pg.LogisticGAM(pg.s(0) + pg.s(1))
        with contextlib.redirect_stdout(io.StringIO()):
            scores = gam.gridsearch(X[:, :2], y, lam=np.logspace(10, 12, 3), return_scores=True)
        assert len(scores) <= 3 and gam._is_fitted
        lin = pg.LinearGAM(pg.s(0) + pg.s(1))
        with contextlib.redirect_stdout(io.StringIO()):
            s2 = lin.gridsearch(X[:, :2], X[:, 0] + y, lam=np.logspace(10, 16, 4), return_scores=True)
        assert 1 <= len(s2) <= 4 and lin._is_fitted


@pytest.mark.parametrize("kind", BACKENDS)
@pytest.mark.parametrize("name", ["c1_wage", "c2_logistic", "gamma", "linear_weights", "poisson_exposure", "tensor_by"])
def test_p_values_standard_errors_aicc(kind, name):
    """statistics_ entries stored in the round-1 fixtures that no test compared"""
    case = cases.FIT_CASES[name]
    g = helpers.load("fit_" + name)
    with backend(kind):
        gam = helpers.fit_case(case, g)
    st = gam.statistics_
    assert abs(st["AICc"] - float(g["stat_AICc"])) <= 1e-9 * abs(float(g["stat_AICc"]))
    Pz = helpers.identifiable_projector(gam)
    if "stat_cov" in g:  # standard errors on the identifiable part (the raw diagonal carries the null-space noise)
        se = np.sqrt(np.maximum(np.diag(Pz @ st["cov"] @ Pz), 0))
        se_ref = np.sqrt(np.maximum(np.diag(Pz @ g["stat_cov"] @ Pz), 0))
        assert rel(se, se_ref) < 1e-6
    # p-values: Wald statistics of the centred spline coefficients; the intercept's cov entry sits on the null space
    pv, pv_ref = np.asarray(st["p_values"], float), g["stat_p_values"]
    assert len(pv) == len(pv_ref) and np.all((pv >= 0) & (pv <= 1))
    # _compute_p_value takes pinv(cov block) with a rank cut-off (pygam.py:1283): scipy's max(M, N) eps in the reference,
    # sqrt(eps) here (cov through the normal equations carries more noise along directions of zero variance: wage's 7
    # years on 20 splines are rank 7 in both).  Blocks where the two cut-offs disagree on the REFERENCE's cov do not compare.
    compared = 0
    for i, t in enumerate(gam.terms):
        if t.isintercept or "stat_cov" not in g:
            continue
        idx = gam.terms.get_coef_indices(i)
        sv = np.linalg.svd(g["stat_cov"][idx][:, idx], compute_uv=False)
        cut = sv.max() * len(idx) * np.finfo(float).eps
        if np.any((sv > cut) & (sv < sv.max() * 1e-7)):
            continue
        compared += 1
        np.testing.assert_allclose(pv[i], pv_ref[i], rtol=1e-5, atol=1e-9)
    assert compared >= (2 if name == "c1_wage" else 1) or "stat_cov" not in g


@pytest.mark.parametrize("kind", BACKENDS)
def test_tensor_terms_with_linear_marginals(kind):
    g = r2("te_linear_marginal")
    with backend(kind):
        gam = pg.PoissonGAM(pg.te(pg.s(0, n_splines=8), pg.l(1), lam=[0.3, 2.0]) + pg.s(2, n_splines=12)
                            + pg.te(pg.l(3), pg.s(2, n_splines=6)))
        quiet_fit(gam, g["X"], g["y"])
        np.testing.assert_array_equal(gam._P(), g["penalty"])
        check_fit(gam, g, g["X"])
        assert rel(gam.predict_mu(g["X"][:64]), g["mu_head"]) < 1e-8
        pd0 = gam.partial_dependence(term=0, X=g["X"][:32])  # the interval pass on a tensor term with a linear marginal
        idx = gam.terms.get_coef_indices(0)
        assert rel(pd0, g["modelmat_head"][:, idx] @ g["coef"][idx]) < 1e-7
    # nested tensors are refused, factor marginals have no device evaluation
    with pytest.raises(ValueError):
        pg.te(pg.l(0), pg.te(0, 1))
    with pytest.raises(NotImplementedError):
        pg.te(pg.s(0), pg.f(1))


@pytest.mark.parametrize("kind", BACKENDS)
def test_fit_quantile(kind):
    g = r2("fit_quantile")
    with backend(kind):
        gam = pg.ExpectileGAM(pg.s(0) + pg.s(1), expectile=0.5)
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gam.fit_quantile(g["X"], g["y"], quantile=0.8)
        assert gam.expectile == float(g["expectile"])  # same bisection path
        assert abs(gam._get_quantile_ratio(g["X"], g["y"]) - float(g["ratio"])) < 1e-12
        assert abs(gam.statistics_["GCV"] - float(g["stat_GCV"])) <= 1e-8 * float(g["stat_GCV"])


@pytest.mark.parametrize("kind", BACKENDS)
def test_pickle_and_deepcopy_of_a_fitted_model(kind):
    g = helpers.load("fit_gamma")
    with backend(kind):
        gam = helpers.fit_case(cases.FIT_CASES["gamma"], g)
        mu = gam.predict_mu(g["X_new"])
        for clone in (pickle.loads(pickle.dumps(gam)), copy.deepcopy(gam)):
            np.testing.assert_array_equal(clone.coef_, gam.coef_)
            assert clone.statistics_["edof"] == gam.statistics_["edof"]
            np.testing.assert_array_equal(clone.predict_mu(g["X_new"]), mu)  # the clone rebuilds its engine
        assert abs(gam.loglikelihood(g["X"], g["y"]) - float(g["stat_loglikelihood"])) <= 1e-9 * abs(float(g["stat_loglikelihood"]))


@pytest.mark.parametrize("kind", BACKENDS)
def test_fewer_rows_than_coefficients(kind):
    """n = 15 rows, m = 41 coefficients: the reference crops its SVD to rank n (pygam.py:789-799,
    pygam/tests/test_GAM_methods.py:93-107); pgb_solve_truncated takes the n largest eigenpairs of G + E^T E"""
    g = r2("rows_lt_coefs")
    with backend(kind):
        gam = pg.LinearGAM(pg.s(0) + pg.s(1))
        quiet_fit(gam, g["X"], g["y"])
        assert len(gam.logs_["diffs"]) == int(g["n_iter"])
        st = gam.statistics_
        assert abs(st["edof"] - float(g["stat_edof"])) < 1e-7
        assert abs(st["GCV"] - float(g["stat_GCV"])) < 1e-7 * float(g["stat_GCV"])
        assert rel(gam._linear_predictor(g["X"]), g["lp_head"]) < 1e-6


@pytest.mark.parametrize("kind", BACKENDS)
def test_config4_grid_11_cubed(kind):
    """BASELINE config 4 at n = 3000: 1331 candidates, one Gram pass, batched solves; scores, iteration counts
    (363 x 1, 968 x 2 on the wage data of the survey; here whatever the reference logged), edof and AIC"""
    g = r2("grid_c4_grid11")
    X, y = cases.data_c4(3000)
    with backend(kind):
        gam = helpers.build_model(cases.FIT_CASES["c4_linear"])
        with contextlib.redirect_stdout(io.StringIO()):
            scores = gam.gridsearch(X, y, return_scores=True, lam=[np.logspace(-3, 3, 11)] * 3)
    models = list(scores)
    assert len(models) == 1331 == len(g["scores"])
    np.testing.assert_allclose(list(scores.values()), g["scores"], rtol=1e-10)
    np.testing.assert_array_equal([len(mm.logs_["diffs"]) for mm in models], g["n_iters"])
    np.testing.assert_allclose([mm.statistics_["edof"] for mm in models], g["edofs"], rtol=1e-9)
    np.testing.assert_allclose([mm.statistics_["AIC"] for mm in models], g["aics"], rtol=1e-9)
    np.testing.assert_allclose([mm.logs_["deviance"][0] for mm in models[1:]], g["first_dev"][1:], rtol=1e-9)
    assert abs(gam.statistics_["GCV"] - float(g["best_score"])) <= 1e-10 * float(g["best_score"])
    np.testing.assert_allclose(np.ravel(gam.lam), g["best_lam"])


def _big(name, gen, n):
    g = r2(name)
    X, y = gen(n)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
    assert sha(X) == str(g["sha_X"]) and sha(y) == str(g["sha_y"]), "the seeded generator drifted from the fixture"
    return g, X, y


@pytest.mark.gpu
def test_config2_at_100k_rows_matches_reference():
    g, X, y = _big("big_c2", cases.data_c2, 100_000)
    gam = helpers.build_model(cases.FIT_CASES["c2_logistic"])
    quiet_fit(gam, X, y)
    check_fit(gam, g, X)


@pytest.mark.gpu
def test_config3_at_50k_rows_matches_reference():
    g, X, y = _big("big_c3", cases.data_c3, 50_000)
    gam = helpers.build_model(cases.FIT_CASES["c3_poisson"])
    quiet_fit(gam, X, y)
    check_fit(gam, g, X)


@pytest.mark.gpu
@pytest.mark.parametrize("m,rank", [(41, 15), (61, 57), (26, 26), (130, 40)])
def test_truncated_solve_kernel_matches_numpy(m, rank):
    """pgb_solve_truncated (Jacobi eigen-solver on the device) vs numpy.linalg.eigh on the same system"""
    import torch
    from oracle_backend import OracleEngine
    from pygam_b200.engine import Engine

    nsp = (m - 1) // 2
    terms = pg.terms.TermList(pg.s(0, n_splines=nsp) + pg.s(1, n_splines=m - 1 - nsp) + pg.terms.Intercept())
    terms.compile(np.random.default_rng(0).random((50, 2)))
    assert terms.n_coefs == m
    eng = Engine(terms, "normal", "identity", n_features=2)
    ora = OracleEngine(terms, "normal", "identity", 1, None, 2, "cpu")
    rng = np.random.default_rng(m)
    B = rng.standard_normal((rank + 3, m))
    G = B.T @ B
    r = B.T @ rng.standard_normal(rank + 3)
    P, _ = terms._penalties(split=True)
    E = np.eye(m) * np.sqrt(np.finfo(float).eps) + 0.6 * P
    for e in (eng, ora):
        e.out[:m * m] = torch.from_numpy(G.ravel()).to(e.out.device)
        e.out[m * m:m * m + m] = torch.from_numpy(r).to(e.out.device)
        e.set_penalty(E)
    prev = rng.standard_normal(m)
    a = eng.solve_truncated(rank, coef_prev=prev, want_cov=True)
    b = ora.solve_truncated(rank, coef_prev=prev, want_cov=True)
    assert a["info"] == 0
    assert rel(a["coef"], b["coef"]) < 1e-9
    assert abs(a["edof"] - b["edof"]) < 1e-9 * max(1.0, abs(b["edof"]))
    assert abs(a["diff"] - b["diff"]) < 1e-9 * b["diff"]
    assert rel(a["cov"], b["cov"]) < 1e-8


@pytest.mark.parametrize("kind", BACKENDS)
def test_non_finite_input_is_refused_by_every_entry_point(kind):
    """check_X / check_y of the reference run wherever external data enters a model (pygam/tests/test_utils.py:108-173):
    NaN in X, y or the weights raises ValueError in fit, predict, the intervals, partial dependence, residuals,
    log-likelihood, score, grid search and sample"""
    rng = np.random.default_rng(3)
    n = 400
    X = rng.uniform(0, 1, (n, 2))
    y = np.sin(5 * X[:, 0]) + X[:, 1] + 0.1 * rng.standard_normal(n)
    w = np.ones(n)
    Xn, yn, wn = X.copy(), y.copy(), w.copy()
    Xn[0, 1], yn[0], wn[0] = np.nan, np.nan, np.nan
    with backend(kind), contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gam = pg.LinearGAM(pg.s(0, n_splines=8) + pg.s(1, n_splines=8))
        for args in ((Xn, y, w), (X, yn, w), (X, y, wn)):
            with pytest.raises(ValueError):
                gam.fit(*args)
        gam.fit(X, y)
        for call in (gam.predict, gam.predict_mu, gam.confidence_intervals, gam.prediction_intervals):
            with pytest.raises(ValueError):
                call(Xn)
        with pytest.raises(ValueError):
            gam.partial_dependence(term=0, X=Xn)
        for call in (gam.deviance_residuals, gam.loglikelihood, gam.score, gam.gridsearch):
            for args in ((Xn, y, w), (X, yn, w), (X, y, wn)):
                with pytest.raises(ValueError):
                    call(*args)
        with pytest.raises(ValueError):
            gam.sample(Xn, y)
        with pytest.raises(ValueError):
            gam.sample(X, yn, weights=w, n_bootstraps=2)
        assert np.isfinite(gam.loglikelihood(X, y, w)) and 0.0 < gam.score(X, y, w) <= 1.0


@pytest.mark.parametrize("kind", BACKENDS)
@pytest.mark.parametrize("name", ["c1_wage", "mixed_terms", "tensor_by"])
def test_term_build_columns_and_params(kind, name):
    """Term.build_columns / TermList.build_columns (terms.py:693-708, 910-957, 1077-1096, 1454-1477, 1901-1923) return
    the scipy-sparse columns of the reference's model matrix (fixture rows); get_params / set_params of a term show
    the parameters the reference shows (core.py:139-192 with the per-kind exclusions of terms.py)"""
    import scipy.sparse

    g = helpers.load("fit_" + name)
    with backend(kind):
        gam = helpers.fit_case(cases.FIT_CASES[name], g)
        B = gam.terms.build_columns(g["X_new"])
        assert scipy.sparse.issparse(B)
        np.testing.assert_allclose(B.toarray(), g["modelmat_new"], rtol=0, atol=5e-13)
        for i, t in enumerate(gam.terms):
            cols = gam.terms.get_coef_indices(i)
            np.testing.assert_allclose(t.build_columns(g["X_new"]).toarray(), g["modelmat_new"][:, cols], rtol=0, atol=5e-13)
            np.testing.assert_array_equal(gam.terms.build_columns(g["X_new"], term=i).toarray(), t.build_columns(g["X_new"]).toarray())
    shown = {"intercept_term": ["verbose"], "linear_term": ["feature", "lam", "penalties", "verbose"],
             "spline_term": ["basis", "by", "constraints", "dtype", "feature", "lam", "n_splines", "penalties", "spline_order", "verbose"],
             "factor_term": ["coding", "feature", "lam", "penalties", "verbose"], "tensor_term": ["by", "verbose"]}
    for t in gam.terms:
        assert sorted(t.get_params()) == shown[t.kind], t
        assert "edge_knots_" in t.get_params(deep=True) or t.kind in ("intercept_term", "tensor_term")
    assert sorted(gam.terms.get_params()) == ["verbose"]
    t = pg.s(0)
    assert t.set_params(n_splines=7, nonsense=1).n_splines == 7 and not hasattr(t, "nonsense")
    assert t.set_params(force=True, nonsense=1).nonsense == 1
