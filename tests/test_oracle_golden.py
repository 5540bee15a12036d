"""Pin the CPU oracle (oracle/gam_oracle.py) to outputs of the unmodified reference.

The fixtures under tests/golden/ were produced by tests/golden/make_golden.py from
/root/reference (pyGAM 0.12.0).  The oracle keeps the reference's QR+SVD method, so on the
same BLAS it agrees to rounding; tolerances below allow for a different BLAS build on the
GPU box: raw coefficients carry the reference's own ~1e-7 null-space noise,
identifiable quantities (lp, mu, deviance, edof, GCV/UBRE) are held to 1e-9 or better.
"""
import os

import numpy as np
import pytest

import cases
from oracle import gam_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("name", sorted(cases.BASIS_CASES))
def test_basis_known_answers(name):
    g = load("basis")
    K, order, periodic, ek = cases.BASIS_CASES[name]
    B = orc.bspline_basis(g[name + "_x"], ek, K, order, periodic)
    assert B.shape == g[name + "_B"].shape
    np.testing.assert_array_equal(B, g[name + "_B"])


def test_tensor_product():
    g = load("basis")
    np.testing.assert_array_equal(orc.tensor_rows(g["tensor_a"], g["tensor_b"]), g["tensor_ab"])


def test_penalties_known_answers():
    g = load("penalties")
    for n in (1, 2, 3, 5, 8, 20):
        coef = g[f"coef_{n}"]
        np.testing.assert_array_equal(orc.pen_derivative(n), g[f"derivative_{n}"])
        np.testing.assert_array_equal(orc.pen_l2(n), g[f"l2_{n}"])
        if n >= 5:
            np.testing.assert_array_equal(orc.pen_periodic(n), g[f"periodic_{n}"])
        for cname in ("monotonic_inc", "monotonic_dec", "convex", "concave"):
            np.testing.assert_array_equal(orc.CONSTRAINTS[cname](n, coef), g[f"{cname}_{n}"])


@pytest.mark.parametrize("name", sorted(cases.FIT_CASES))
def test_fit_matches_reference(name):
    case = cases.FIT_CASES[name]
    g = load("fit_" + name)
    res = orc.fit(case, g["X"], g["y"], weights=g.get("weights"), exposure=g.get("exposure"))
    st = res["statistics"]
    # model matrix and penalty: exact
    mm = res["modelmat"][:24].toarray()
    np.testing.assert_array_equal(mm, g["modelmat_head"])
    np.testing.assert_array_equal(orc.penalty_matrix(res["terms"]), g["penalty"])
    np.testing.assert_array_equal(orc.model_matrix(res["terms"], g["X_new"]).toarray(), g["modelmat_new"])
    # iteration count, convergence, logs
    assert len(res["logs"]["diffs"]) == int(g["n_iter"])
    assert res["converged"] == bool(g["log_diffs"][-1] < g["tol"])
    np.testing.assert_allclose(res["logs"]["deviance"], g["log_deviance"], rtol=1e-9)
    if "log_accuracy" in g:
        np.testing.assert_allclose(res["logs"]["accuracy"], g["log_accuracy"], rtol=0, atol=1e-12)
    # identifiable quantities
    lp = res["modelmat"][:64].dot(res["coef"])
    assert rel(lp, g["lp_head"]) < 1e-8
    assert rel(orc.predict_mu(res, g["X_new"]), g["mu_new"]) < 1e-8
    for key, tol in [("edof", 1e-9), ("deviance", 1e-9), ("scale", 1e-9), ("loglikelihood", 1e-9),
                     ("AIC", 1e-9), ("AICc", 1e-9)]:
        assert abs(st[key] - float(g["stat_" + key])) <= tol * max(1.0, abs(float(g["stat_" + key]))), key
    for key in ("GCV", "UBRE"):
        ref = float(g["stat_" + key])
        if np.isnan(ref):
            assert st[key] is None
        else:
            assert abs(st[key] - ref) <= 1e-10 * abs(ref), key
    for k, v in st["pseudo_r2"].items():
        assert abs(v - float(g["stat_r2_" + k])) < 1e-9
    # raw coefficients: reference's own noise floor
    assert rel(res["coef"], g["coef"]) < 1e-5
    if "constraint_final" in g:
        C = orc.constraint_matrix(res["terms"], g["coef"], orc.CONSTRAINT_LAM, res["constraint_l2"])
        np.testing.assert_allclose(C, g["constraint_final"], rtol=1e-15, atol=0)


def test_config1_published_known_answers():
    """docs/notebooks/quick_start.ipynb:113-118 and tour_of_pygam.ipynb:375-380"""
    w = load("wage_xy")
    res = orc.fit(cases.FIT_CASES["c1_wage"], w["X"], w["y"])
    st = res["statistics"]
    assert round(st["edof"], 4) == 25.1911
    assert round(st["GCV"], 4) == 1255.6902
    assert round(st["pseudo_r2"]["explained_deviance"], 4) == 0.2955
    assert len(res["logs"]["diffs"]) == 2
    gs = orc.gridsearch(cases.FIT_CASES["c1_wage"], w["X"], w["y"])
    best = gs["results"][gs["best"]]
    assert round(gs["candidates"][gs["best"]], 4) == 15.8489
    assert round(best["statistics"]["edof"], 4) == 19.2602
    assert round(best["statistics"]["GCV"], 4) == 1250.3656


@pytest.mark.parametrize("name", sorted(cases.GRID_CASES))
def test_gridsearch_matches_reference(name):
    gcase = cases.GRID_CASES[name]
    case = cases.FIT_CASES[gcase["fit"]]
    g = load("grid_" + name)
    f = load("fit_" + gcase["fit"])
    grid = None if gcase["grid"] is None else gcase["grid"]["lam"]
    gs = orc.gridsearch(case, f["X"], f["y"], weights=f.get("weights"), exposure=f.get("exposure"),
                        lam_grid=grid)
    assert len(gs["scores"]) == len(g["scores"])
    np.testing.assert_allclose(gs["scores"], g["scores"], rtol=1e-9)
    np.testing.assert_array_equal([len(r["logs"]["diffs"]) for r in gs["results"]], g["n_iters"])
    np.testing.assert_allclose([r["statistics"]["edof"] for r in gs["results"]], g["edofs"], rtol=1e-8)
    assert abs(gs["scores"][gs["best"]] - float(g["best_score"])) <= 1e-10 * abs(float(g["best_score"]))
