"""Generate tests/golden/intervals.npz by running the UNMODIFIED reference (authoring container
only, like make_golden.py): confidence / prediction intervals and partial dependence
(pygam.py:1297-1643, 2477-2506) of a few fitted parity cases.

    python tests/golden/make_golden_intervals.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402
import make_golden as mg  # noqa: E402

CASES = ["c1_wage", "c2_logistic", "gamma", "tensor3", "mixed_terms", "tensor_by", "linear_weights"]


def main():
    pg = mg._import_reference()
    wz = np.load(os.path.join(HERE, "wage_xy.npz"))
    wage_xy = (wz["X"], wz["y"])
    out = {}
    for name in CASES:
        case = cases.FIT_CASES[name]
        data = cases.make_data(case["data"], wage_xy)
        weights = cases.make_weights(case.get("weights"), len(data["y"]))
        gam = mg.build_model(pg, case)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mg.fit_model(gam, case, data, weights)
        rng = np.random.default_rng(77)
        Xn = data["X"][rng.choice(len(data["X"]), size=40, replace=False)]
        out[f"{name}__X"] = Xn
        out[f"{name}__ci"] = gam.confidence_intervals(Xn, width=0.95)
        out[f"{name}__ci_q"] = gam.confidence_intervals(Xn, quantiles=[0.1, 0.5, 0.9])
        if case["model"] == "LinearGAM":
            out[f"{name}__pi"] = gam.prediction_intervals(Xn, width=0.9)
        for i, term in enumerate(gam.terms):
            if term.isintercept:
                continue
            out[f"{name}__grid{i}"] = gam.generate_X_grid(term=i, n=12)
            pdep, conf = gam.partial_dependence(term=i, X=gam.generate_X_grid(term=i, n=12), width=0.95)
            out[f"{name}__pdep{i}"] = pdep
            out[f"{name}__pconf{i}"] = conf
            if term.istensor:
                XX = gam.generate_X_grid(term=i, n=7, meshgrid=True)
                pm, cm = gam.partial_dependence(term=i, X=XX, quantiles=[0.25, 0.75], meshgrid=True)
                out[f"{name}__mesh_pdep{i}"] = pm
                out[f"{name}__mesh_conf{i}"] = cm
        out[f"{name}__pdep_only"] = gam.partial_dependence(term=0, X=gam.generate_X_grid(term=0, n=12))
    np.savez_compressed(os.path.join(HERE, "intervals.npz"), **out)
    print("wrote intervals.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
