"""The plan of the cell-owned Gram pass for models with te() / f() / l() terms, checked on the CPU: the host-side
self check (pygam_b200/csrc/pgb_selfcheck.cu, libpgb_selfcheck.so -- test infrastructure) walks the rows through the
same plan, record / key / moment arithmetic, change of basis and assembly as the kernels of pgb_cells.cu, and its
[G | r] must equal the oracle's X^T W^2 X, X^T W^2 z on the same inputs."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import cases
import helpers
from oracle import gam_oracle as orc
from oracle_backend import OracleBackend, OracleEngine
from pygam_b200 import _lib
from pygam_b200 import engine as pg_engine
from pygam_b200.engine import DeviceData

SELF = os.path.join(os.path.dirname(_lib.LIB_PATH), "libpgb_selfcheck.so")
_P = C.c_void_p


def _selfcheck_lib():
    lib = C.CDLL(SELF)
    fn = lib.pgb_selfcheck_cells_gram
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(_lib.PgbTerm), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double,
                   C.POINTER(_lib.PgbDebugOpts), C.c_int32, _P, C.c_int64, C.c_int64, _P, _P, _P, _P, C.POINTER(C.c_int32)]
    return fn


def _fit_on_cpu(case, g):
    old = pg_engine.swap_backend_for_tests(OracleBackend())
    try:
        gam = helpers.fit_case(case, g, max_iter=2)
    finally:
        pg_engine.swap_backend_for_tests(old)
    return gam


def _check(case, g, tile_rows=0, threads=0, tol=2e-13):
    fn = _selfcheck_lib()
    gam = _fit_on_cpu(case, g)
    F = g["X"].shape[1]
    y, w = orc.prepare_weights(case["model"], g["y"], g.get("weights"), g.get("exposure"))
    has_w = not (g.get("weights") is None and g.get("exposure") is None)
    w32 = np.ascontiguousarray(w, dtype=np.float32) if has_w else None
    Xt = np.ascontiguousarray(g["X"].T, dtype=np.float64)
    n = Xt.shape[1]
    ora = OracleEngine(gam.terms, gam.distribution.name, gam.link.name, gam.distribution.levels, gam._expectile, F, "cpu")
    data = DeviceData(torch.from_numpy(Xt), torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64)),
                      None if w32 is None else torch.from_numpy(w32))
    m = gam.terms.n_coefs
    rng = np.random.default_rng(1)
    coef = np.ascontiguousarray(gam.coef_ * (1 + 0.05 * rng.standard_normal(m)))
    desc = pg_engine.build_descriptor(gam.terms, F)
    fam = ora.fam
    opts = _lib.PgbDebugOpts(0, tile_rows, 0, threads)
    ngen = C.c_int32(0)
    for mode in (_lib.MODE_IRLS, _lib.MODE_INIT):
        if mode == _lib.MODE_INIT and case["model"] == "LinearGAM":
            continue
        out = np.zeros(m * m + m)
        yy = np.ascontiguousarray(y, dtype=np.float64)
        rc = fn(desc, len(gam.terms), F, _lib.DISTS[gam.distribution.name], _lib.LINKS[gam.link.name],
                float(gam.distribution.levels), float(gam._expectile) if gam._expectile is not None else -1.0,
                C.byref(opts), mode, Xt.ctypes.data, n, n, yy.ctypes.data, None if w32 is None else w32.ctypes.data,
                coef.ctypes.data, out.ctypes.data, C.byref(ngen))
        assert rc == 0, rc
        ref = ora.gram_pass(data, coef, mode=mode).numpy()
        G, Gr = out[:m * m].reshape(m, m), ref[:m * m].reshape(m, m)
        scale = np.abs(Gr).max()
        assert np.abs(G - Gr).max() <= tol * scale, (mode, np.abs(G - Gr).max() / scale)
        r, rr = out[m * m:], ref[m * m:m * m + m]
        assert np.abs(r - rr).max() <= tol * max(np.abs(rr).max(), scale)
    del fam
    return ngen.value


@pytest.mark.parametrize("name,min_generic", [("c1_wage", 5), ("c3_poisson", 40), ("dummy_coding", 3),
                                              ("multi_penalty", 5), ("c2_logistic", 0), ("linear_weights", 0),
                                              ("edge_knots", 0)])
def test_cell_plan_reproduces_the_oracle_gram(name, min_generic):
    case = cases.FIT_CASES[name]
    g = helpers.load("fit_" + name)
    assert _check(case, g) >= min_generic


@pytest.mark.parametrize("tile_rows,threads", [(256, 256), (1024, 320), (64, 512)])
def test_cell_plan_tiles_and_key_ranges(tile_rows, threads):
    """small tiles (hash splits depend on the row's position in its tile) and narrow CTAs (more key ranges)"""
    case = cases.FIT_CASES["c1_wage"]
    _check(case, helpers.load("fit_c1_wage"), tile_rows=tile_rows, threads=threads)


def test_cell_plan_te_f_l_and_rows_outside_the_knots():
    """te() x f() x l() x s() in one model, te() twice (4-d cells), rows outside the compiled edge knots"""
    rng = np.random.default_rng(7)
    n = 1500
    X = np.column_stack([rng.random(n), rng.random(n), rng.integers(0, 4, n).astype(float), rng.standard_normal(n),
                         rng.random(n), rng.random(n)])
    eta = np.sin(4 * X[:, 0]) * X[:, 1] + 0.3 * X[:, 2] + 0.2 * X[:, 3] + X[:, 4] * X[:, 5]
    y = rng.poisson(np.exp(0.4 * eta)).astype(float)
    s, te, f, l = cases.s, cases.te, cases.f, cases.l
    case = dict(model="PoissonGAM",
                terms=[te(s(0, n_splines=6), s(1, n_splines=5)), f(2), l(3), s(4, n_splines=7, edge_knots=[0.1, 0.8]),
                       te(s(5, n_splines=5, edge_knots=[0.2, 0.9]), s(0, n_splines=4))])
    assert _check(case, dict(X=X, y=y), tol=5e-13) >= 15


def test_cell_plan_by_variables_and_linear_marginals():
    """s(j, by=k) and te(s, l) / te(l, s): a cubic factor times a value factor, as generic pairs of the cell-owned pass
    (against each other, against te(s, s), f(), l() and plain s(); rows outside the knots; weights)"""
    rng = np.random.default_rng(11)
    n = 1800
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
    assert _check(case, dict(X=X, y=y, weights=w), tol=5e-13) >= 30
