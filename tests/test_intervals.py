"""Confidence / prediction intervals and partial dependence (pygam.py:1297-1643,
2477-2506) against golden vectors of the unmodified reference (tests/golden/intervals.npz, made by
tests/golden/make_golden_intervals.py).  The host logic runs on CPU with the oracle-backed test
engine; the GPU variant calls pgb_interval_pass through the C ABI."""
import numpy as np
import pytest

import cases
import helpers
from helpers import fit_case, load
from oracle_backend import OracleBackend
from pygam_b200 import engine as pg_engine

CASES = ["c1_wage", "c2_logistic", "gamma", "tensor3", "mixed_terms", "tensor_by", "linear_weights"]
# the interval half-widths inherit the covariance's sensitivity to the near-null directions of the
# design: reference vs itself differs at 1e-7 there
RTOL = 2e-6
# the LEVEL of a single term's partial dependence is not identifiable (a constant can move between
# the term and the intercept; the reference's own raw coefficients carry ~1e-7 relative noise along
# that direction): shapes are compared tightly, levels loosely
LEVEL_TOL = 2e-5


# absolute floor on the linear-predictor scale (O(1) in every case): a heavily penalised term whose
# whole contribution is ~1e-7 has no meaningful relative accuracy in either implementation
ABS_FLOOR = 2e-6


def _assert_pdep(pdep, conf, gp, gc):
    scale = max(np.abs(gp).max(), 1e-12)
    shift = np.mean(pdep - gp)
    assert abs(shift) <= LEVEL_TOL * scale + ABS_FLOOR
    np.testing.assert_allclose(pdep - shift, gp, rtol=RTOL, atol=2e-7 * scale + ABS_FLOOR)
    if conf is not None:
        np.testing.assert_allclose(conf - pdep[..., None], gc - gp[..., None], rtol=5e-6, atol=2e-7 * scale + ABS_FLOOR)


def _check(name):
    gold = load("intervals")
    case = cases.FIT_CASES[name]
    g = load("fit_" + name)
    gam = fit_case(case, g)
    Xn = gold[f"{name}__X"]
    np.testing.assert_allclose(gam.confidence_intervals(Xn, width=0.95), gold[f"{name}__ci"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(gam.confidence_intervals(Xn, quantiles=[0.1, 0.5, 0.9]), gold[f"{name}__ci_q"], rtol=RTOL,
                               atol=1e-9)
    if case["model"] == "LinearGAM":
        np.testing.assert_allclose(gam.prediction_intervals(Xn, width=0.9), gold[f"{name}__pi"], rtol=RTOL, atol=1e-9)
    for i, term in enumerate(gam.terms):
        if term.isintercept:
            with pytest.raises(ValueError):
                gam.partial_dependence(term=i)
            continue
        grid = gam.generate_X_grid(term=i, n=12)
        np.testing.assert_allclose(grid, gold[f"{name}__grid{i}"], rtol=1e-14, atol=0)
        pdep, conf = gam.partial_dependence(term=i, X=grid, width=0.95)
        _assert_pdep(pdep, conf, gold[f"{name}__pdep{i}"], gold[f"{name}__pconf{i}"])
        if term.istensor:
            XX = gam.generate_X_grid(term=i, n=7, meshgrid=True)
            pm, cm = gam.partial_dependence(term=i, X=XX, quantiles=[0.25, 0.75], meshgrid=True)
            assert pm.shape == gold[f"{name}__mesh_pdep{i}"].shape and cm.shape == gold[f"{name}__mesh_conf{i}"].shape
            _assert_pdep(pm, cm, gold[f"{name}__mesh_pdep{i}"], gold[f"{name}__mesh_conf{i}"])
    only = gam.partial_dependence(term=0, X=gam.generate_X_grid(term=0, n=12))
    _assert_pdep(only, None, gold[f"{name}__pdep_only"], None)
    with pytest.raises(ValueError):
        gam.confidence_intervals(Xn, quantiles=[0.0, 0.5])
    with pytest.raises(ValueError):
        gam.partial_dependence(term=len(gam.terms))


@pytest.mark.parametrize("name", CASES)
def test_intervals_host_logic(name):
    old = pg_engine.swap_backend_for_tests(OracleBackend())
    try:
        _check(name)
    finally:
        pg_engine.swap_backend_for_tests(old)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_intervals_gpu(name):
    _check(name)
