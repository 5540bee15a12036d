"""Host-side planning of the Gram pass (no GPU): for every model structure of the parity cases
the plan must produce each entry of the upper triangle of [G | r] in exactly one shared-memory
accumulator element, and must fit the B200's 227 KB of shared memory per CTA."""
import ctypes as C

import numpy as np
import pytest

import cases
import helpers
from oracle_backend import OracleBackend
from pygam_b200 import _lib
from pygam_b200 import engine as pg_engine


def plan_for(case, X):
    old = pg_engine.swap_backend_for_tests(OracleBackend())
    try:
        gam = helpers.build_model(case)
        data = gam._prepare_data(X, None, None, fitting=False)
        mm, _ = pg_engine.column_stats(data)
        gam._validate_data_dep_params(data, mm)
    finally:
        pg_engine.swap_backend_for_tests(old)
    desc = pg_engine.build_descriptor(gam.terms, X.shape[1])
    info = _lib.PgbModelInfo()
    errs = C.c_int32(-1)
    rc = _lib.load().pgb_model_plan_host(desc, len(gam.terms), X.shape[1], C.byref(info), C.byref(errs))
    _lib.check(rc, "pgb_model_plan_host")
    return gam, info, errs.value


@pytest.mark.parametrize("name", sorted(cases.FIT_CASES))
def test_plan_covers_upper_triangle_exactly_once(name):
    g = helpers.load("fit_" + name)
    gam, info, errs = plan_for(cases.FIT_CASES[name], g["X"])
    assert errs == 0
    assert info.m == gam.terms.n_coefs
    assert info.n_roles >= 1 and info.tile_rows >= 16 and info.gram_ctas >= info.n_roles


def test_plan_of_baseline_configs():
    """config 2 (m=251, k=41) and config 3 (m=601, k=113) need several roles"""
    rng = np.random.default_rng(0)
    _, info2, e2 = plan_for(cases.FIT_CASES["c2_logistic"], rng.random((50, 10)))
    assert (info2.m, info2.k, e2) == (251, 41, 0)
    assert info2.n_roles >= 2
    _, info3, e3 = plan_for(cases.FIT_CASES["c3_poisson"], rng.random((50, 24)))
    assert (info3.m, info3.k, e3) == (601, 113, 0)
    assert info3.n_roles >= 8
