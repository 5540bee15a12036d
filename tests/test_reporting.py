"""The reporting layer over a fitted model: summary() (pygam.py:1645-1806), edof_per_coef
(pygam.py:1033), deviance_residuals (940-991), score (917-938), sample and its bootstrap refits (2087-2368).
Fixtures: the reference's own output (tests/golden/r2_reporting.npz, make_golden_r2.py reporting)."""
import contextlib
import io
import os
import warnings

import numpy as np
import pytest

import cases
import helpers
from test_parity_gaps import BACKENDS, backend
import pygam_b200 as pg

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _summary(gam):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gam.summary()
    return buf.getvalue()


@pytest.mark.parametrize("kind", BACKENDS)
def test_summary_prints_the_reference_table(kind):
    g = dict(np.load(os.path.join(GOLD, "r2_reporting.npz")))
    with backend(kind):
        for name in ("poisson_exposure", "dummy_coding"):  # designs whose per-term edof the reference pins down
            gam = helpers.fit_case(cases.FIT_CASES[name], helpers.load("fit_" + name))
            assert _summary(gam) == str(g["summary_" + name]), name
        for name in ("c1_wage", "gamma"):
            # null directions: the EDoF column (edof_per_coef) carries the reference's rounding noise;
            # every other cell of the table is identical
            gam = helpers.fit_case(cases.FIT_CASES[name], helpers.load("fit_" + name))
            mine, ref = _summary(gam).splitlines(), str(g["summary_" + name]).splitlines()
            assert len(mine) == len(ref)
            for a, b in zip(mine, ref):
                if a != b:  # a term row: every column but EDoF must still agree (feature function, lambda, rank, P > x, sig. code)
                    ta, tb = a.split(), b.split()
                    assert a[:67] == b[:67] and len(ta) == len(tb) and ta[:3] == tb[:3] and ta[4:] == tb[4:], (a, b)
                    assert abs(float(ta[3]) - float(tb[3])) <= 0.25, (a, b)
        with pytest.raises(AttributeError):
            pg.LinearGAM().summary()


@pytest.mark.parametrize("kind", BACKENDS)
def test_edof_per_coef(kind):
    with backend(kind):
        for name in ("tensor_by", "poisson_exposure", "dummy_coding", "no_intercept"):
            g = helpers.load("fit_" + name)
            gam = helpers.fit_case(cases.FIT_CASES[name], g)
            np.testing.assert_allclose(gam.statistics_["edof_per_coef"], g["stat_edof_per_coef"], rtol=0, atol=1e-8)
        for name in ("c2_logistic", "gamma", "c5_expectile"):
            # the split along the null directions is the reference's rounding noise; the total is not
            g = helpers.load("fit_" + name)
            gam = helpers.fit_case(cases.FIT_CASES[name], g)
            epc = gam.statistics_["edof_per_coef"]
            assert abs(epc.sum() - gam.statistics_["edof"]) < 1e-6
            assert np.abs(epc - g["stat_edof_per_coef"]).max() < 0.1


@pytest.mark.parametrize("kind", BACKENDS)
def test_deviance_residuals_and_score(kind):
    g = dict(np.load(os.path.join(GOLD, "r2_reporting.npz")))
    f = helpers.load("fit_gamma")
    with backend(kind):
        gam = helpers.fit_case(cases.FIT_CASES["gamma"], f)
        np.testing.assert_allclose(gam.deviance_residuals(f["X"], f["y"]), g["gamma_dev_resid"], rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(gam.deviance_residuals(f["X"], f["y"], scaled=True), g["gamma_dev_resid_scaled"],
                                   rtol=1e-7, atol=1e-10)
        assert abs(gam.score(f["X"], f["y"]) - float(g["gamma_score"])) < 1e-9
        p = helpers.load("fit_poisson_exposure")
        pgam = helpers.fit_case(cases.FIT_CASES["poisson_exposure"], p)
        assert abs(pgam.score(p["X"], p["y"] / p["exposure"], weights=p["exposure"]) - float(g["poisson_score"])) < 1e-9


@pytest.mark.parametrize("kind", BACKENDS)
def test_sample_shapes_and_moments(kind):
    """pygam/tests/test_GAM_methods.py:463-521 checks shapes; here also that coefficient draws centre on coef_"""
    f = helpers.load("fit_c4_linear")
    with backend(kind), contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gam = helpers.fit_case(cases.FIT_CASES["c4_linear"], f)
        np.random.seed(0)
        X, y = f["X"][:500], f["y"][:500]
        coef = gam.sample(X, y, quantity="coef", n_draws=400, n_bootstraps=1)
        assert coef.shape == (400, len(gam.coef_))
        Pz = helpers.identifiable_projector(gam)
        se = np.sqrt(np.maximum(np.diag(Pz @ gam.statistics_["cov"] @ Pz), 1e-12))
        assert np.all(np.abs((coef.mean(0) - gam.coef_) @ Pz) < 6 * se / np.sqrt(400) + 1e-6)
        mu = gam.sample(X, y, quantity="mu", n_draws=7, n_bootstraps=2, sample_at_X=X[:30])
        assert mu.shape == (7, 30)
        ys = gam.sample(X, y, quantity="y", n_draws=5, n_bootstraps=1)
        assert ys.shape == (5, 500)
        with pytest.raises(ValueError):
            gam.sample(X, y, quantity="foo")
        with pytest.raises(ValueError):
            gam.sample(X, y, n_bootstraps=0)
