"""cProfile of LogisticGAM.fit from numpy arrays (config 2 shape): where the host time of a fit goes.
python tools/fitprofile.py [--rows 10000000]"""
import argparse
import contextlib
import cProfile
import io
import os
import pstats
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402
import pygam_b200 as pg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--sync", action="store_true", help="CUDA_LAUNCH_BLOCKING-like attribution: synchronise after every engine call")
args = ap.parse_args()
X, y = cases.data_c2(args.rows)
X = np.ascontiguousarray(X)
terms = None
for j in range(X.shape[1]):
    t = pg.s(j, n_splines=25)
    terms = t if terms is None else terms + t
with contextlib.redirect_stdout(io.StringIO()):
    pg.LogisticGAM(terms).fit(X[:200_000], y[:200_000])
    gam = pg.LogisticGAM(terms)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    gam.fit(X, y)
    torch.cuda.synchronize()
    pr.disable()
st = pstats.Stats(pr, stream=sys.stdout)
st.sort_stats("cumulative").print_stats(45)
st.sort_stats("tottime").print_stats(25)
