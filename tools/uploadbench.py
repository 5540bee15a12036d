"""Host rows -> device shard: pgb_upload_rows against a pageable torch copy + device transpose (what from_host did
before), and the wall time of LogisticGAM.fit from numpy arrays.   python tools/uploadbench.py [--rows 10000000]"""
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
from pygam_b200 import _lib, engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    args = ap.parse_args()
    n = args.rows
    X, y = cases.data_c2(n)
    X = np.ascontiguousarray(X, dtype=np.float64)
    F = X.shape[1]
    dev = torch.device("cuda:0")
    lib = _lib.load()

    def old():
        buf = torch.zeros((F, n), dtype=torch.float64, device=dev)
        buf.copy_(torch.from_numpy(X).to(dev).T)
        yt = torch.from_numpy(y).to(dev)
        torch.cuda.synchronize()
        return buf, yt

    def new(threads):
        buf = torch.empty((F, n), dtype=torch.float64, device=dev)
        yt = torch.empty(n, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        _lib.check(lib.pgb_upload_rows(buf.data_ptr(), n, X.ctypes.data, n, F, threads))
        _lib.check(lib.pgb_upload_rows(yt.data_ptr(), n, y.ctypes.data, n, 1, threads))
        return buf, yt

    for name, fn in [("pageable torch copy + device transpose", old)] + [(f"pgb_upload_rows, {t} threads", lambda t=t: new(t)) for t in (1, 2, 4, 8, 12, 16)]:
        fn()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        b = min(ts)
        print(f"{name}: {b * 1e3:.1f} ms  ({(n * (F + 1) * 8) / b / 1e9:.1f} GB/s)", flush=True)
    terms = None
    for j in range(F):
        t = pg.s(j, n_splines=25)
        terms = t if terms is None else terms + t
    with contextlib.redirect_stdout(io.StringIO()):
        pg.LogisticGAM(terms).fit(X[:200_000], y[:200_000])
        for _ in range(3):
            gam = pg.LogisticGAM(terms)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            gam.fit(X, y)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(f"LogisticGAM.fit from numpy, {n} rows: {dt * 1e3:.1f} ms, {len(gam.logs_['diffs'])} iterations", file=sys.stderr, flush=True)


if __name__ == "__main__":
    main()
