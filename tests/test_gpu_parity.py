"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, is
compared with the CPU oracle on the same seeded inputs and with the golden fixtures produced by
the unmodified reference.  Tolerances are the north star's: identifiable coefficients 1e-8,
deviance / GCV / UBRE 1e-10 relative, identical iteration counts and convergence outcome.
Nothing here reads /root/reference."""
import ctypes as C

import numpy as np
import pytest
import torch

import cases
import helpers
from helpers import check_against_golden, fit_case, load
from oracle import gam_oracle as orc
from oracle_backend import OracleBackend, OracleEngine, solve_mirror, householder
import pygam_b200 as pg
from pygam_b200 import _lib
from pygam_b200.engine import DeviceData, Engine

pytestmark = pytest.mark.gpu


def _cpu_data(g, case):
    y, w = orc.prepare_weights(case["model"], g["y"], g.get("weights"), g.get("exposure"))
    wt = None if (g.get("weights") is None and g.get("exposure") is None) else torch.from_numpy(
        np.asarray(w, dtype=np.float32))
    return DeviceData(torch.from_numpy(np.ascontiguousarray(g["X"].T)), torch.from_numpy(y), wt)


def _compiled(case, g):
    """product model with compiled terms (fit on the GPU) + matching CPU-side oracle engine"""
    gam = fit_case(case, g)
    eng = gam._engine
    ora = OracleEngine(gam.terms, gam.distribution.name, gam.link.name, gam.distribution.levels,
                       gam._expectile, g["X"].shape[1], "cpu")
    return gam, eng, ora


# ---------------------------------------------------------------------------------------------
# basis: model-matrix kernel vs the reference's b_spline_basis known answers
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(cases.BASIS_CASES))
def test_basis_known_answers(name):
    g = load("basis")
    K, order, periodic, ek = cases.BASIS_CASES[name]
    x = g[name + "_x"]
    term = pg.s(0, n_splines=K, spline_order=order, basis="cp" if periodic else "ps", edge_knots=list(ek))
    tl = pg.terms.TermList(term)
    tl.compile(np.zeros((1, 1)))
    eng = Engine(tl, "normal", "identity", n_features=1)
    data = DeviceData(torch.from_numpy(x[None, :].copy()).cuda())
    B = eng.model_matrix(data).cpu().numpy()
    ref = g[name + "_B"]
    assert B.shape == ref.shape
    np.testing.assert_allclose(B, ref, rtol=2e-13, atol=2e-14)


@pytest.mark.parametrize("name", sorted(cases.FIT_CASES))
def test_model_matrix_matches_reference(name):
    case = cases.FIT_CASES[name]
    g = load("fit_" + name)
    gam = fit_case(case, g)
    np.testing.assert_allclose(gam._modelmat(g["X"][:24]), g["modelmat_head"], rtol=0, atol=5e-14)
    np.testing.assert_allclose(gam._modelmat(g["X_new"]), g["modelmat_new"], rtol=0, atol=5e-13)
    np.testing.assert_array_equal(gam._P(), g["penalty"])


# ---------------------------------------------------------------------------------------------
# kernels vs oracle on the same inputs
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1_wage", "c2_logistic", "c3_poisson", "c5_expectile", "gamma", "invgauss",
                                  "logistic_weights", "poisson_exposure", "mixed_terms", "tensor3", "tensor_by",
                                  "dummy_coding", "orders", "gam_generic"])
def test_gram_pass_matches_oracle(name):
    case = cases.FIT_CASES[name]
    g = load("fit_" + name)
    gam, eng, ora = _compiled(case, g)
    m = eng.m
    cpu = _cpu_data(g, case)
    gpu = DeviceData(cpu.Xt.cuda(), cpu.y.cuda(), None if cpu.w is None else cpu.w.cuda())
    rng = np.random.default_rng(1)
    coef = g["coef"] * (1 + 0.05 * rng.standard_normal(m))
    modes = [_lib.MODE_IRLS] + ([] if case["model"] == "LinearGAM" else [_lib.MODE_INIT])
    for mode in modes:
        out = eng.gram_pass(gpu, coef, mode=mode).cpu().numpy()
        ref = ora.gram_pass(cpu, coef, mode=mode).numpy()
        G, Gr = out[:m * m].reshape(m, m), ref[:m * m].reshape(m, m)
        scale = np.abs(Gr).max()
        assert np.abs(G - Gr).max() <= 1e-12 * scale
        np.testing.assert_array_equal(G, G.T)
        r, rr = out[m * m:m * m + m], ref[m * m:m * m + m]
        assert np.abs(r - rr).max() <= 1e-12 * max(np.abs(rr).max(), scale)
        sc, scr = out[m * m + m:], ref[m * m + m:]
        assert sc[_lib.GS_NUSED] == scr[_lib.GS_NUSED]
        assert sc[_lib.GS_NROWS] == scr[_lib.GS_NROWS]
        assert sc[_lib.GS_NCORRECT] == scr[_lib.GS_NCORRECT]
        assert abs(sc[_lib.GS_DEV] - scr[_lib.GS_DEV]) <= 1e-12 * max(1.0, abs(scr[_lib.GS_DEV]))
    # bit-reproducible: partial sums are combined in a fixed order
    a = eng.gram_pass(gpu, coef).clone()
    b = eng.gram_pass(gpu, coef).clone()
    assert torch.equal(a, b)


@pytest.mark.parametrize("name", ["c2_logistic", "c3_poisson", "gamma", "invgauss", "linear_weights",
                                  "poisson_exposure", "c1_wage"])
def test_stats_pass_matches_oracle(name):
    case = cases.FIT_CASES[name]
    g = load("fit_" + name)
    gam, eng, ora = _compiled(case, g)
    cpu = _cpu_data(g, case)
    gpu = DeviceData(cpu.Xt.cuda(), cpu.y.cuda(), None if cpu.w is None else cpu.w.cuda())
    scale = float(g["stat_scale"])
    ybar = float(cpu.y.mean())
    pr = case["model"] == "PoissonGAM"
    sc, lp, mu = eng.stats_pass(gpu, g["coef"], scale=scale, ybar=ybar, poisson_round=pr, want_mu=True, want_lp=True)
    scr, lpr, mur = ora.stats_pass(cpu, g["coef"], scale=scale, ybar=ybar, poisson_round=pr, want_mu=True, want_lp=True)
    assert helpers.rel(lp.cpu().numpy(), lpr.numpy()) < 1e-12
    assert helpers.rel(mu.cpu().numpy(), mur.numpy()) < 1e-12
    for k in range(_lib.SS_NONFINITE):
        assert abs(sc[k] - scr[k]) <= 1e-11 * max(1.0, abs(scr[k])), (k, sc[k], scr[k])


@pytest.mark.parametrize("name", ["c2_logistic", "c3_poisson", "c5_expectile"])
def test_stats_pass_extrapolated_rows_and_unaligned_inputs(name):
    """rows outside the edge knots (the tangent-line entries of the per-interval polynomial table, utils.py:744-763)
    and inputs the TMA ring cannot take (odd leading dimension: plain loads) against the oracle"""
    case = cases.FIT_CASES[name]
    g = load("fit_" + name)
    gam, eng, ora = _compiled(case, g)
    cpu = _cpu_data(g, case)
    X = cpu.Xt.numpy().T.copy()
    lo, hi = X.min(axis=0), X.max(axis=0)
    X = lo + (X - lo) * 1.3 - 0.15 * (hi - lo)  # ~23 % of every feature outside [lo, hi], both sides
    X[0, :] = lo  # exactly on the edge knots
    X[1, :] = hi
    n = X.shape[0] - (1 - X.shape[0] % 2)  # odd number of rows
    for rows, padded in ((X.shape[0], True), (n, False)):
        c = DeviceData.from_host(X[:rows], cpu.y.numpy()[:rows], None if cpu.w is None else cpu.w.numpy()[:rows], device="cpu")
        if padded:
            gpu = DeviceData.from_host(X[:rows], cpu.y.numpy()[:rows], None if cpu.w is None else cpu.w.numpy()[:rows])
            assert gpu.ld % 2 == 0
        else:
            gpu = DeviceData(torch.from_numpy(np.ascontiguousarray(X[:rows].T)).cuda(), c.y.cuda(), None if c.w is None else c.w.cuda())
            assert gpu.ld % 2 == 1
        sc, lp, mu = eng.stats_pass(gpu, g["coef"], want_mu=True, want_lp=True)
        scr, lpr, mur = ora.stats_pass(c, g["coef"], want_mu=True, want_lp=True)
        assert helpers.rel(lp.cpu().numpy(), lpr.numpy()) < 1e-12
        fin = np.isfinite(mur.numpy())
        assert helpers.rel(mu.cpu().numpy()[fin], mur.numpy()[fin]) < 1e-11
        for k in (_lib.SS_N, _lib.SS_SUMY):
            assert abs(sc[k] - scr[k]) <= 1e-11 * max(1.0, abs(scr[k])), (k, sc[k], scr[k])


def test_stats_pass_more_terms_than_the_plan_holds():
    """52 s() terms (the plan of the statistics pass holds 48; the rest take the general evaluation after the table
    loop), 10 te() terms (the plan holds 8), a factor and a linear term; rows outside the edge knots; vs the oracle"""
    rng = np.random.default_rng(5)
    n, F = 3000, 74
    X = rng.uniform(0.0, 1.0, size=(n, F))
    X[:, 72] = rng.integers(0, 4, size=n)
    y = np.sin(3 * X[:, 0]) + X[:, 1] ** 2 + 0.1 * rng.standard_normal(n)
    terms = pg.s(0, n_splines=8)
    for j in range(1, 52):
        terms = terms + pg.s(j, n_splines=6 + j % 5)
    for j in range(10):
        terms = terms + pg.te(52 + 2 * j, 53 + 2 * j, n_splines=5)
    terms = terms + pg.f(72) + pg.l(73)
    gam = pg.LinearGAM(terms, max_iter=1)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        gam.fit(X, y)
    eng = gam._engine
    ora = OracleEngine(gam.terms, gam.distribution.name, gam.link.name, gam.distribution.levels, gam._expectile, F, "cpu")
    coef = rng.standard_normal(eng.m)
    Xn = X.copy()
    Xn[:, :72] = Xn[:, :72] * 1.2 - 0.1  # ~17 % of every continuous feature outside the edge knots
    c = DeviceData.from_host(Xn, y, device="cpu")
    g = DeviceData.from_host(Xn, y)
    sc, lp, mu = eng.stats_pass(g, coef, want_mu=True, want_lp=True)
    scr, lpr, mur = ora.stats_pass(c, coef, want_mu=True, want_lp=True)
    assert helpers.rel(lp.cpu().numpy(), lpr.numpy()) < 1e-12
    assert helpers.rel(mu.cpu().numpy(), mur.numpy()) < 1e-12
    for k in (_lib.SS_N, _lib.SS_DEV, _lib.SS_PEARSON, _lib.SS_SUMY):
        assert abs(sc[k] - scr[k]) <= 1e-11 * max(1.0, abs(scr[k])), (k, sc[k], scr[k])


@pytest.mark.parametrize("n_terms,n_splines", [(3, 150), (7, 150), (12, 40)])
def test_stats_pass_large_interval_tables_and_many_tiles(n_terms, n_splines):
    """the interval tables of the pass in eight bank-disjoint copies (3 x 150 splines: 113 KB, 12 x 40: 114 KB) and in one
    copy when eight do not fit beside the staged tiles (7 x 150 splines); 150 001 rows = several tiles of the TMA ring per
    CTA, a ragged tail, rows outside the edge knots; statistics and predict instances, with and without weights"""
    rng = np.random.default_rng(17)
    n = 150_001
    X = rng.uniform(0.0, 1.0, size=(n, n_terms))
    eta = np.sin(5 * X[:, 0]) + np.cos(3 * X[:, 1]) * X[:, 2]
    y = rng.poisson(np.exp(0.4 * eta)).astype(float)
    terms = pg.s(0, n_splines=n_splines)
    for j in range(1, n_terms):
        terms = terms + pg.s(j, n_splines=n_splines - j)
    gam = pg.PoissonGAM(terms, max_iter=1)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        gam.fit(X[:4000], y[:4000])
    eng = gam._engine
    ora = OracleEngine(gam.terms, gam.distribution.name, gam.link.name, gam.distribution.levels, gam._expectile, n_terms, "cpu")
    coef = 0.3 * rng.standard_normal(eng.m)
    Xn = X * 1.1 - 0.05  # ~9 % of every feature outside the edge knots
    w = rng.uniform(0.5, 2.0, n).astype(np.float32)
    for weights in (None, w):
        c = DeviceData.from_host(Xn, y, weights, device="cpu")
        g = DeviceData.from_host(Xn, y, weights)
        sc, lp, mu = eng.stats_pass(g, coef, want_mu=True, want_lp=True)
        scr, lpr, mur = ora.stats_pass(c, coef, want_mu=True, want_lp=True)
        assert helpers.rel(lp.cpu().numpy(), lpr.numpy()) < 1e-12
        assert helpers.rel(mu.cpu().numpy(), mur.numpy()) < 1e-12
        for k in (_lib.SS_N, _lib.SS_DEV, _lib.SS_WDEV, _lib.SS_PEARSON, _lib.SS_SUMY, _lib.SS_SUMW):
            assert abs(sc[k] - scr[k]) <= 1e-11 * max(1.0, abs(scr[k])), (k, sc[k], scr[k])
        _, _, mu_only = eng.stats_pass(DeviceData(g.Xt, None, None), coef, want_mu=True)  # the predict instance (no y)
        assert helpers.rel(mu_only.cpu().numpy(), mur.numpy()) < 1e-12
        # the IRLS step of a prepared Gram pass runs through the same kernel: against the oracle's Gram matrix
        first = eng.gram_pass(g, coef).cpu().numpy().copy()
        planned = eng.gram_pass(g, coef).cpu().numpy().copy()
        ref = ora.gram_pass(c, coef).numpy()
        m = eng.m
        scale = np.abs(ref[:m * m]).max()
        assert np.abs(first - ref)[:m * m + m].max() <= 1e-12 * scale
        assert np.abs(planned - ref)[:m * m + m].max() <= 1e-12 * scale
        for k in (_lib.GS_NUSED, _lib.GS_NROWS):
            assert planned[m * m + m + k] == ref[m * m + m + k]
        assert abs(planned[m * m + m + _lib.GS_DEV] - ref[m * m + m + _lib.GS_DEV]) <= 1e-11 * abs(ref[m * m + m + _lib.GS_DEV])
    # the solve of this system against the numpy mirror: m = 1030 (7 x 150 splines) is past the 687 rows of the panel
    # that pgb_solve stages in shared memory (the rest is read from the matrix) and past the 860 of pgb_chol_check
    m = eng.m
    out = ora.gram_pass(DeviceData.from_host(Xn, y, None, device="cpu"), coef).numpy().copy()
    Pn, Pr = gam.terms._penalties(split=True)
    EtE = np.eye(m) * np.sqrt(orc.EPS) + Pr
    eng.out.copy_(torch.from_numpy(out))
    eng.set_penalty(EtE, Pn)
    sol = eng.solve(coef_prev=coef, want_edof=True, want_cov=True)
    V, tau = householder(gam.terms.null_vectors())
    ref = solve_mirror(out[:m * m].reshape(m, m), out[m * m:m * m + m], EtE, V, tau, want_cov=True, EtE_null=Pn)
    assert sol["info"] == 0
    Pz = helpers.identifiable_projector(gam)
    assert np.linalg.norm(Pz @ (sol["coef"] - ref["coef"])) <= 1e-9 * np.linalg.norm(ref["coef"])
    assert abs(sol["edof"] - ref["edof"]) <= 1e-9 * abs(ref["edof"])
    assert helpers.rel(Pz @ sol["cov"] @ Pz, Pz @ ref["cov"] @ Pz) < 1e-7
    assert eng.chol_ok(EtE + Pn)
    assert not eng.chol_ok(EtE + Pn - 5.0 * np.eye(m))


@pytest.mark.parametrize("name", ["c1_wage", "c2_logistic", "c3_poisson", "mixed_terms", "no_intercept",
                                  "c5_expectile"])
def test_solve_matches_numpy_mirror(name):
    case = cases.FIT_CASES[name]
    g = load("fit_" + name)
    gam, eng, ora = _compiled(case, g)
    m = eng.m
    cpu = _cpu_data(g, case)
    out = ora.gram_pass(cpu, g["coef"]).numpy().copy()
    Pn, Pr = gam.terms._penalties(split=True)
    np.testing.assert_array_equal(Pn + Pr, gam._P())
    if gam.terms.hasconstraint:  # 1e9-stiff constraint penalty, all of it null-annihilating
        Cn, Cr = gam.terms._constraints(g["coef"], gam._constraint_lam, gam._constraint_l2, split=True)
        Pn, Pr = Pn + Cn, Pr + Cr
        assert np.abs(Cn).max() >= 1e9
    EtE = np.eye(m) * np.sqrt(orc.EPS) + Pr
    eng.out.copy_(torch.from_numpy(out))
    eng.set_penalty(EtE, Pn)
    sol = eng.solve(coef_prev=g["coef"], want_edof=True, want_cov=True)
    V, tau = householder(gam.terms.null_vectors())
    ref = solve_mirror(out[:m * m].reshape(m, m), out[m * m:m * m + m], EtE, V, tau, want_cov=True, EtE_null=Pn)
    assert sol["info"] == 0
    Pz = helpers.identifiable_projector(gam)
    # the 1e9-stiff constrained system (cond ~1e13) sits at the rounding limit in any solver
    stiff = gam.terms.hasconstraint
    assert np.linalg.norm(Pz @ (sol["coef"] - ref["coef"])) <= (1e-6 if stiff else 1e-10) * np.linalg.norm(ref["coef"])
    assert np.linalg.norm(sol["coef"] - ref["coef"]) <= (1e-6 if stiff else 1e-8) * np.linalg.norm(ref["coef"])
    assert abs(sol["edof"] - ref["edof"]) <= (1e-7 if stiff else 1e-10) * abs(ref["edof"])
    assert helpers.rel(Pz @ sol["cov"] @ Pz, Pz @ ref["cov"] @ Pz) < (1e-4 if stiff else 1e-8)
    d = np.linalg.norm(g["coef"] - sol["coef"]) / np.linalg.norm(sol["coef"])
    assert abs(sol["diff"] - d) <= 1e-9 * max(d, 1e-300) + 1e-18
    # not positive definite -> info 1, non-finite input -> info 2 (pygam.py:785-786)
    eng.set_penalty(-np.eye(m) * 1e12)
    assert eng.solve()["info"] == 1
    bad = out.copy()
    bad[3] = np.nan
    eng.out.copy_(torch.from_numpy(bad))
    eng.set_penalty(EtE)
    assert eng.solve()["info"] == 2
    assert eng.chol_ok(EtE)
    assert not eng.chol_ok(EtE - 5.0 * np.eye(m))


def test_host_buffer_gram_equals_device_gram():
    """pgb_gram_pass_host (chunked H2D staging on two streams) == pgb_gram_pass"""
    case = cases.FIT_CASES["c2_logistic"]
    X, y = cases.data_c2(50_000, seed=5)
    gam = helpers.build_model(case).fit(X, y)
    eng = gam._engine
    m = eng.m
    data = DeviceData.from_host(X, y)
    ref = eng.gram_pass(data, gam.coef_).cpu().numpy().copy()
    Xt = torch.from_numpy(np.ascontiguousarray(X.T)).pin_memory()
    yt = torch.from_numpy(y).pin_memory()
    beta = np.ascontiguousarray(gam.coef_)
    out = np.empty(m * m + m + _lib.GS_COUNT)
    for chunk in (0, 7_000, 50_000):
        rc = eng.lib.pgb_gram_pass_host(eng.handle, _lib.MODE_IRLS, C.c_void_p(Xt.data_ptr()), len(y), len(y),
                                        C.c_void_p(yt.data_ptr()), C.c_void_p(0), beta.ctypes.data_as(C.c_void_p),
                                        out.ctypes.data_as(C.c_void_p), chunk)
        _lib.check(rc, "pgb_gram_pass_host")
        G, Gr = out[:m * m], ref[:m * m]
        assert np.abs(G - Gr).max() <= 1e-12 * np.abs(Gr).max()
        np.testing.assert_allclose(out[m * m:m * m + m], ref[m * m:m * m + m], rtol=1e-10, atol=1e-9)
        assert out[m * m + m + _lib.GS_NUSED] == ref[m * m + m + _lib.GS_NUSED]


# ---------------------------------------------------------------------------------------------
# full fits and grid searches vs the reference's golden fixtures
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(cases.FIT_CASES))
def test_fit_matches_reference(name):
    case = cases.FIT_CASES[name]
    g = load("fit_" + name)
    gam = fit_case(case, g)
    check_against_golden(gam, g, constrained=name in helpers.CONSTRAINED)


@pytest.mark.parametrize("name", sorted(cases.GRID_CASES))
def test_gridsearch_matches_reference(name):
    gcase = cases.GRID_CASES[name]
    case = cases.FIT_CASES[gcase["fit"]]
    g = load("grid_" + name)
    f = load("fit_" + gcase["fit"])
    gam = helpers.build_model(case)
    kw = {} if gcase["grid"] is None else dict(gcase["grid"])
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


# ---------------------------------------------------------------------------------------------
# size-independent properties at sizes the oracle cannot reach
# ---------------------------------------------------------------------------------------------
def _device_c2(n, seed):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    Xt = torch.rand((10, n), dtype=torch.float64, device="cuda", generator=gen)
    eta = torch.zeros(n, dtype=torch.float64, device="cuda")
    for j in range(10):
        eta += 0.5 * torch.sin(2 * np.pi * (j + 1) * Xt[j] / 3.0)
    y = (torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) < torch.sigmoid(eta)).double()
    return DeviceData(Xt, y)


def test_stats_pass_sums_match_written_mu_and_add_over_row_shards():
    """4e6 rows of config 2: the sums of the statistics pass equal the same sums taken (torch, fp64) from the mu the
    predict pass writes; shards with a ragged tail and an odd leading dimension (plain loads) add up to the whole"""
    n = 4_000_000
    data = _device_c2(n, 91)
    gam = helpers.build_model(cases.FIT_CASES["c2_logistic"], max_iter=2)
    gam.fit(data)
    eng = gam._engine
    coef = gam.coef_
    sc, _, _ = eng.stats_pass(data, coef, reduce=False)
    _, lp, mu = eng.stats_pass(DeviceData(data.Xt), coef, want_mu=True, want_lp=True, reduce=False)
    y = data.y
    assert sc[_lib.SS_N] == n and sc[_lib.SS_NONFINITE] == 0
    assert abs(sc[_lib.SS_SUMY] - float(y.sum())) <= 1e-12 * n
    assert float((mu - torch.sigmoid(lp)).abs().max()) < 1e-15 * 8
    dev = float((-2.0 * torch.log(torch.where(y == 1.0, mu, 1.0 - mu))).sum())
    pearson = float(((y - mu) ** 2 / (mu * (1.0 - mu))).sum())
    assert abs(sc[_lib.SS_DEV] - dev) <= 1e-12 * dev
    assert abs(sc[_lib.SS_WDEV] - dev) <= 1e-12 * dev
    assert abs(sc[_lib.SS_PEARSON] - pearson) <= 1e-12 * pearson
    assert sc[_lib.SS_NCORRECT] == float((y == (mu > 0.5).double()).sum())
    h = 1_234_567  # odd: the second shard starts 8-byte aligned only
    a, _, _ = eng.stats_pass(DeviceData(data.Xt[:, :h].contiguous(), y[:h].contiguous()), coef, reduce=False)
    b, _, _ = eng.stats_pass(DeviceData(data.Xt[:, h:].contiguous(), y[h:].contiguous()), coef, reduce=False)
    for k in (_lib.SS_N, _lib.SS_NCORRECT, _lib.SS_SUMY):
        assert a[k] + b[k] == sc[k]
    for k in (_lib.SS_DEV, _lib.SS_PEARSON):
        assert abs(a[k] + b[k] - sc[k]) <= 1e-12 * abs(sc[k])
    # bit-reproducible run to run
    sc2, _, _ = eng.stats_pass(data, coef, reduce=False)
    assert np.array_equal(sc, sc2)


def test_gram_is_additive_over_row_shards_and_oracle_checked_on_a_prefix():
    """G, r and the scalars of 2e6 rows equal the sum over two row shards (what the all-reduce
    relies on), and the Gram of a 20k-row prefix equals the oracle's"""
    n = 2_000_000
    data = _device_c2(n, 77)
    gam = helpers.build_model(cases.FIT_CASES["c2_logistic"], max_iter=2)
    gam.fit(data)
    eng = gam._engine
    m = eng.m
    coef = gam.coef_
    full = eng.gram_pass(data, coef).clone()
    h = 1_234_567
    a = eng.gram_pass(DeviceData(data.Xt[:, :h].contiguous(), data.y[:h].contiguous()), coef).clone()
    b = eng.gram_pass(DeviceData(data.Xt[:, h:].contiguous(), data.y[h:].contiguous()), coef).clone()
    s = a + b
    scale = full[:m * m].abs().max()
    assert (full[:m * m + m] - s[:m * m + m]).abs().max() <= 1e-12 * scale
    assert float(full[m * m + m + _lib.GS_NUSED]) == float(s[m * m + m + _lib.GS_NUSED])
    # partition of unity: every spline block of a row sums to one, so each block of G sums to
    # the total weight  (sum_ij G[block_a, block_b] = sum_i omega_i)
    G = full[:m * m].view(m, m).cpu().numpy()
    tot = G[-1, -1]
    for blk in range(10):
        sl = slice(25 * blk, 25 * blk + 25)
        assert abs(G[sl, sl].sum() - tot) <= 1e-10 * tot
        assert abs(G[sl, -1].sum() - tot) <= 1e-10 * tot
    # prefix vs oracle
    k = 20_000
    cpu = DeviceData(data.Xt[:, :k].cpu().contiguous(), data.y[:k].cpu().contiguous())
    ora = OracleEngine(gam.terms, "binomial", "logit", 1, None, 10, "cpu")
    ref = ora.gram_pass(cpu, coef).numpy()
    got = eng.gram_pass(DeviceData(data.Xt[:, :k].contiguous(), data.y[:k].contiguous()), coef).cpu().numpy()
    assert np.abs(got[:m * m] - ref[:m * m]).max() <= 1e-12 * np.abs(ref[:m * m]).max()


def test_large_fit_matches_oracle_on_subsample_statistics():
    """2e5-row LogisticGAM fit on the GPU (what the O(n m^2) oracle finishes in seconds; the reference itself at 1e5 /
    5e4 rows: tests/test_parity_gaps.py big_c2 / big_c3): the same number of iterations as the oracle on the same rows,
    UBRE / edof / deviance to 1e-10"""
    n = 200_000
    data = _device_c2(n, 123)
    gam = helpers.build_model(cases.FIT_CASES["c2_logistic"]).fit(data)
    X = data.Xt.cpu().numpy().T.copy()
    y = data.y.cpu().numpy()
    ref = orc.fit(cases.FIT_CASES["c2_logistic"], X, y)
    assert len(gam.logs_["diffs"]) == len(ref["logs"]["diffs"])
    for key in ("edof", "UBRE", "deviance", "loglikelihood"):
        a, b = gam.statistics_[key], ref["statistics"][key]
        assert abs(a - b) <= 1e-10 * abs(b), (key, a, b)
    np.testing.assert_allclose(gam.logs_["deviance"], ref["logs"]["deviance"], rtol=1e-10)
