"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run only in the authoring container, where the reference is mounted read-only:

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference (``/root/reference``, pyGAM 0.12.0) imports ``progressbar`` at
``pygam/pygam.py:9``; that package is not installed, so a 3-line stub module is put on
``sys.path`` first.  Nothing else of the reference is altered.  The fixtures hold the
inputs (X, y, weights, exposure) and the reference outputs, so the GPU box -- which has no
``/root/reference`` -- can check parity against them.
"""

import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

REFERENCE = os.environ.get("PYGAM_REFERENCE", "/root/reference")


def _import_reference():
    stub = tempfile.mkdtemp(prefix="pgstub_")
    with open(os.path.join(stub, "progressbar.py"), "w") as fh:
        fh.write("class ProgressBar:\n    def __call__(self, it):\n        return it\n")
    sys.path.insert(0, stub)
    sys.path.insert(0, REFERENCE)
    import pygam  # noqa: F401

    return pygam


def build_term(pg, d):
    d = dict(d)
    kind = d.pop("kind")
    if kind == "s":
        return pg.s(d.pop("feature"), **d)
    if kind == "l":
        return pg.l(d.pop("feature"), **d)
    if kind == "f":
        return pg.f(d.pop("feature"), **d)
    if kind == "te":
        margs = [build_term(pg, m) for m in d.pop("marginals")]
        return pg.te(*margs, **d)
    raise KeyError(kind)


def build_model(pg, case):
    terms = None
    for t in case["terms"]:
        bt = build_term(pg, t)
        terms = bt if terms is None else terms + bt
    cls = getattr(pg, case["model"])
    return cls(terms, **case.get("model_kwargs", {}))


def fit_model(gam, case, data, weights):
    if case["model"] == "PoissonGAM":
        return gam.fit(data["X"], data["y"], exposure=data.get("exposure"), weights=weights)
    return gam.fit(data["X"], data["y"], weights=weights)


def x_new(X, terms, rng):
    """rows for predict(): in-range rows plus numeric columns pushed outside the edge knots"""
    idx = rng.choice(len(X), size=48, replace=False)
    Xn = X[idx].copy()
    cat = set()
    for t in terms:
        if t["kind"] == "f":
            cat.add(t["feature"])
    for j in range(X.shape[1]):
        if j in cat:
            continue
        lo, hi = X[:, j].min(), X[:, j].max()
        w = hi - lo
        Xn[:8, j] = lo - rng.random(8) * 0.3 * w
        Xn[8:16, j] = hi + rng.random(8) * 0.3 * w
    return Xn


def stats_dict(gam):
    st = gam.statistics_
    out = {}
    for k in ["edof", "scale", "GCV", "UBRE", "AIC", "AICc", "loglikelihood", "deviance",
              "n_samples", "m_features"]:
        v = st.get(k)
        out["stat_" + k] = np.nan if v is None else np.float64(v)
    if len(gam.coef_) <= 150:  # keep the fixtures small
        out["stat_cov"] = np.asarray(st["cov"])
    out["stat_se"] = np.asarray(st["se"])
    out["stat_edof_per_coef"] = np.asarray(st["edof_per_coef"])
    out["stat_p_values"] = np.asarray(st["p_values"], dtype=np.float64)
    for k, v in st["pseudo_r2"].items():
        out["stat_r2_" + k] = np.float64(v)
    return out


def run_fit_case(pg, name, case, wage_xy):
    import cases

    data = cases.make_data(case["data"], wage_xy)
    n = len(data["y"])
    weights = cases.make_weights(case.get("weights"), n)
    gam = build_model(pg, case)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fit_model(gam, case, data, weights)
    out = dict(X=data["X"], y=data["y"])
    if weights is not None:
        out["weights"] = weights
    if "exposure" in data:
        out["exposure"] = data["exposure"]
    out["coef"] = np.asarray(gam.coef_)
    for k, v in gam.logs_.items():
        out["log_" + k] = np.asarray(v, dtype=np.float64)
    out.update(stats_dict(gam))
    out["n_iter"] = np.int64(len(gam.logs_["diffs"]))
    out["tol"] = np.float64(gam.tol)
    out["lp_head"] = gam._linear_predictor(data["X"][:64])
    out["mu_head"] = gam.predict_mu(data["X"][:64])
    rng = np.random.default_rng(2024)
    Xn = x_new(data["X"], case["terms"], rng)
    out["X_new"] = Xn
    out["mu_new"] = gam.predict_mu(Xn)
    out["modelmat_head"] = gam._modelmat(data["X"][:24]).toarray()
    out["modelmat_new"] = gam._modelmat(Xn).toarray()
    out["edge_knots"] = np.asarray(pg.utils.flatten(gam.terms.edge_knots_), dtype=np.float64)
    out["penalty"] = gam._P().toarray()
    if gam.terms.hasconstraint:
        out["constraint_final"] = gam._C().toarray()
    return gam, out


def run_grid_case(pg, name, gcase, fit_cases, wage_xy):
    import cases

    case = fit_cases[gcase["fit"]]
    data = cases.make_data(case["data"], wage_xy)
    weights = cases.make_weights(case.get("weights"), len(data["y"]))
    gam = build_model(pg, case)
    kwargs = {} if gcase["grid"] is None else dict(gcase["grid"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if case["model"] == "PoissonGAM":
            scores = gam.gridsearch(data["X"], data["y"], exposure=data.get("exposure"),
                                    weights=weights, return_scores=True, progress=False, **kwargs)
        else:
            scores = gam.gridsearch(data["X"], data["y"], weights=weights, return_scores=True,
                                    progress=False, **kwargs)
    models = list(scores.keys())
    out = dict(
        scores=np.asarray(list(scores.values()), dtype=np.float64),
        lams=np.asarray([np.ravel(pg.utils.flatten(m.lam)) for m in models], dtype=np.float64),
        n_iters=np.asarray([len(m.logs_["diffs"]) for m in models], dtype=np.int64),
        edofs=np.asarray([m.statistics_["edof"] for m in models], dtype=np.float64),
        best_coef=np.asarray(gam.coef_),
        best_lam=np.ravel(pg.utils.flatten(gam.lam)).astype(np.float64),
        best_edof=np.float64(gam.statistics_["edof"]),
        best_score=np.float64(min(scores.values())),
    )
    return out


def run_basis_cases(pg):
    import cases

    out = {}
    for name, (K, order, periodic, ek) in cases.BASIS_CASES.items():
        x = cases.basis_x(ek, periodic)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            B = pg.utils.b_spline_basis(x, np.asarray(ek, dtype=np.float64), n_splines=K,
                                        spline_order=order, sparse=False, periodic=periodic,
                                        verbose=False)
        out[name + "_x"] = x
        out[name + "_B"] = np.asarray(B, dtype=np.float64)
    a = np.random.default_rng(0).random((5, 3))
    b = np.random.default_rng(1).random((5, 4))
    out["tensor_a"], out["tensor_b"] = a, b
    out["tensor_ab"] = pg.utils.tensor_product(a, b)
    return out


def run_penalty_cases(pg):
    out = {}
    pen = pg.penalties
    rng = np.random.default_rng(4)
    for n in (1, 2, 3, 5, 8, 20):
        coef = rng.standard_normal(n)
        out[f"coef_{n}"] = coef
        out[f"derivative_{n}"] = pen.derivative(n, None).toarray()
        out[f"l2_{n}"] = pen.l2(n, None).toarray()
        if n >= 5:
            out[f"periodic_{n}"] = pen.periodic(n, None).toarray()
        for cname in ("monotonic_inc", "monotonic_dec", "convex", "concave"):
            out[f"{cname}_{n}"] = np.asarray(getattr(pen, cname)(n, coef).toarray())
    return out


def main():
    pg = _import_reference()
    import cases
    from pygam.datasets import wage

    wage_xy = wage()
    np.savez_compressed(os.path.join(HERE, "wage_xy.npz"), X=wage_xy[0], y=wage_xy[1])
    np.savez_compressed(os.path.join(HERE, "basis.npz"), **run_basis_cases(pg))
    np.savez_compressed(os.path.join(HERE, "penalties.npz"), **run_penalty_cases(pg))
    for name, case in cases.FIT_CASES.items():
        gam, out = run_fit_case(pg, name, case, wage_xy)
        np.savez_compressed(os.path.join(HERE, f"fit_{name}.npz"), **out)
        print(f"fit {name}: n={len(out['y'])} m={len(out['coef'])} iters={int(out['n_iter'])} "
              f"edof={float(out['stat_edof']):.6f}")
    for name, gcase in cases.GRID_CASES.items():
        out = run_grid_case(pg, name, gcase, cases.FIT_CASES, wage_xy)
        np.savez_compressed(os.path.join(HERE, f"grid_{name}.npz"), **out)
        print(f"grid {name}: {len(out['scores'])} models, best={float(out['best_score']):.6f}")


if __name__ == "__main__":
    main()
