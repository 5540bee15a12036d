"""Model shapes beyond the BASELINE configs, fitted on the device and compared with the CPU oracle's fit of the same
data: many terms (m = 1201: past the 687 panel rows pgb_solve stages in shared memory), a 25 x 25 tensor term (its
625 x 625 block of G does not fit the general Gram path: the cell-owned pass alone carries the model), many l() / f()
terms, a mix of everything the general path takes, a tiny sample -- and the one shape neither Gram path takes."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases as cs  # noqa: E402
import helpers  # noqa: E402
from oracle import gam_oracle as orc  # noqa: E402
from pygam_b200 import _lib  # noqa: E402

pytestmark = pytest.mark.gpu


def _data(name, rng, n):
    if name == "many_terms":
        X = rng.random((n, 60))
        eta = np.sin(6 * X[:, 0]) + X[:, 1] ** 2 - X[:, 2]
        return X, (rng.random(n) < 1 / (1 + np.exp(-eta))).astype(float)
    if name == "big_tensor":
        X = rng.random((n, 3))
        return X, rng.poisson(np.exp(0.5 * np.sin(5 * X[:, 0]) * X[:, 1] + 0.3 * X[:, 2])).astype(float)
    if name == "linear_and_factor":
        X = rng.standard_normal((n, 35))
        X[:, 30:] = rng.integers(0, 6, (n, 5))
        return X, X[:, :30] @ rng.standard_normal(30) * 0.1 + X[:, 31] * 0.2 + 0.3 * rng.standard_normal(n)
    if name == "gamma_mixed":
        X = rng.random((n, 6))
        mu = np.exp(0.3 * np.sin(2 * np.pi * X[:, 0]) + 0.2 * X[:, 1] * X[:, 2] + 0.1 * X[:, 3])
        return X, rng.gamma(5.0, mu / 5.0)
    X = rng.random((40, 1))
    return X, np.sin(6 * X[:, 0]) + 0.1 * rng.standard_normal(40)


CASES = {
    "many_terms": {"model": "LogisticGAM", "terms": [cs.s(j, n_splines=20) for j in range(60)]},
    "big_tensor": {"model": "PoissonGAM", "terms": [cs.te(cs.s(0, n_splines=25), cs.s(1, n_splines=25)), cs.s(2)]},
    "linear_and_factor": {"model": "LinearGAM", "terms": [cs.l(j) for j in range(30)] + [cs.f(j) for j in range(30, 35)]},
    "gamma_mixed": {"model": "GammaGAM", "terms": [cs.s(0, basis="cp", n_splines=12), cs.s(1, by=2, n_splines=9),
                                                  cs.s(3, spline_order=2, n_splines=8),
                                                  cs.te(cs.s(3, n_splines=5), cs.s(4, n_splines=6), cs.s(5, n_splines=4))]},
    "tiny_sample": {"model": "LinearGAM", "terms": [cs.s(0)]},
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_fit_of_unusual_shapes_matches_the_oracle(name):
    rng = np.random.default_rng(7)
    X, y = _data(name, rng, 12000)
    case = CASES[name]
    with contextlib.redirect_stdout(io.StringIO()):
        gam = helpers.build_model(case).fit(X, y)
    ref = orc.fit(case, X, y)
    assert len(gam.logs_["diffs"]) == len(ref["logs"]["diffs"])
    st, rst = gam.statistics_, ref["statistics"]
    # gamma_mixed: feature 3 sits in an s() term and in the tensor term (concurvity): directions outside the analytic
    # null space that only the penalty pins; fitted values and deviance agree, raw coefficients to 1e-6
    loose = name == "gamma_mixed"
    assert abs(st["edof"] - rst["edof"]) <= (1e-7 if loose else 1e-9) * rst["edof"]
    assert abs(st["deviance"] - rst["deviance"]) <= 1e-10 * abs(rst["deviance"])
    Pz = helpers.identifiable_projector(gam)
    assert np.linalg.norm(Pz @ (gam.coef_ - ref["coef"])) <= (1e-6 if loose else 1e-8) * np.linalg.norm(ref["coef"])
    mu = gam.predict_mu(X[:500])
    assert np.max(np.abs(mu - orc.predict_mu(ref, X[:500]))) <= 1e-8 * max(1.0, np.max(np.abs(mu)))


def test_a_basis_neither_gram_path_takes_fails_loudly():
    """s(n_splines=200): more than the 127 knot cells per marginal the cell records hold, and its 200 x 200 block of G
    does not fit one SM's shared memory on the general path -- a clear error, no fallback"""
    rng = np.random.default_rng(8)
    X = rng.random((3000, 2))
    y = np.sin(20 * X[:, 0]) + X[:, 1] + 0.1 * rng.standard_normal(3000)
    gam = helpers.build_model({"model": "LinearGAM", "terms": [cs.s(0, n_splines=200), cs.s(1, n_splines=150)]})
    with pytest.raises(_lib.PgbError, match="does not fit in shared memory"):
        with contextlib.redirect_stdout(io.StringIO()):
            gam.fit(X, y)
