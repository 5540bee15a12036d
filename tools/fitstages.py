"""Wall time of the stages of LogisticGAM.fit from numpy arrays (config 2 shape), synchronised around every stage.
python tools/fitstages.py [--rows 10000000] [--blas-before]"""
import argparse
import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402
import pygam_b200 as pg  # noqa: E402
from pygam_b200 import engine, gam as gam_mod  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--blas-before", action="store_true", help="a threaded BLAS call right before every fit")
ap.add_argument("--no-sync", action="store_true")
args = ap.parse_args()
X, y = cases.data_c2(args.rows)
X = np.ascontiguousarray(X)
T = {}


def timed(obj, name, label=None):
    fn = getattr(obj, name)
    label = label or name

    def wrap(*a, **k):
        if not args.no_sync:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(*a, **k)
        if not args.no_sync:
            torch.cuda.synchronize()
        T[label] = T.get(label, 0.0) + time.perf_counter() - t0
        return r

    setattr(obj, name, wrap)


timed(engine.DeviceData, "from_host")
timed(gam_mod, "vector_range")
timed(gam_mod, "column_stats")
timed(gam_mod.GAM, "_validate_data_dep_params")
timed(gam_mod.GAM, "_get_engine")
timed(gam_mod.GAM, "_initial_estimate")
timed(gam_mod.GAM, "_estimate_model_statistics")
timed(gam_mod.GAM, "_pirls")
timed(gam_mod.GAM, "_prepare_data")
terms = None
for j in range(X.shape[1]):
    t = pg.s(j, n_splines=25)
    terms = t if terms is None else terms + t
A = np.random.rand(1500, 1500)
with contextlib.redirect_stdout(io.StringIO()):
    pg.LogisticGAM(terms).fit(X[:200_000], y[:200_000])
    for rep in range(4):
        T.clear()
        gam = pg.LogisticGAM(terms)
        if args.blas_before:
            A @ A
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gam.fit(X, y)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"fit {dt * 1e3:.1f} ms: " + ", ".join(f"{k} {v * 1e3:.1f}" for k, v in T.items()), file=sys.stderr, flush=True)
