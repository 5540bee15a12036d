"""Kernel micro-benchmark on one GPU: times the Gram pass, the statistics pass and the solve of a
parity-case model at n synthetic rows (CUDA events, after warm-up).  Development tool."""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402
import helpers  # noqa: E402
from pygam_b200 import _lib  # noqa: E402
from pygam_b200.engine import DeviceData  # noqa: E402

GEN = {"c2_logistic": (cases.data_c2, 10), "c3_poisson": (cases.data_c3, 24), "c4_linear": (cases.data_c4, 3),
       "c5_expectile": (cases.data_c5, 4)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="c2_logistic")
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only-gram", action="store_true")
    ap.add_argument("--profile", action="store_true", help="cudaProfilerStart/Stop around one Gram pass of the prepared "
                    "shard (ncu --profile-from-start off)")
    a = ap.parse_args()
    gen, F = GEN[a.case]
    Xs, ys = gen(20000)
    case = cases.FIT_CASES[a.case]
    import contextlib, io
    gam = helpers.build_model(case, max_iter=2)
    with contextlib.redirect_stdout(io.StringIO()):
        gam.fit(Xs, ys)
    eng = gam._engine
    # tile the small sample up to n rows on the device
    reps = (a.rows + len(ys) - 1) // len(ys)
    Xt = torch.from_numpy(np.ascontiguousarray(Xs.T)).cuda().repeat(1, reps)[:, :a.rows].contiguous()
    y = torch.from_numpy(ys).cuda().repeat(reps)[:a.rows].contiguous()
    data = DeviceData(Xt, y)
    lib = _lib.load()
    coef = gam.coef_
    _lib.check(lib.pgb_gram_timing(eng.handle, 1), "timing")
    eng.REFINE_AFTER_PASSES = 1 << 40  # the second preparation step is run (and timed) explicitly below
    path = (f"cells roles={eng.info.cell_roles} tile={eng.info.cell_tile_rows} threads={eng.info.cell_threads} "
            f"ring_groups={eng.info.cell_ring_groups}" if eng.info.cell_roles else
            f"roles={eng.info.n_roles} tile={eng.info.tile_rows} ctas={eng.info.gram_ctas}")

    def measure(tag):
        ms, rs, ts = [], [], []
        for _ in range(a.reps + 2):
            eng.gram_pass(data, coef)
            f, r, t = C.c_float(), C.c_float(), C.c_float()
            _lib.check(lib.pgb_gram_last_ms(eng.handle, C.byref(r), C.byref(f), C.byref(t)), "ms")
            ms.append(f.value)
            rs.append(r.value)
            ts.append(t.value)
        g, r, t = float(np.median(ms[2:])), float(np.median(rs[2:])), float(np.median(ts[2:]))
        print(f"{a.case} n={a.rows} m={eng.m} {path} [{tag}]: "
              f"records {r:.3f} ms | accum {g:.3f} ms | pass {t:.3f} ms = {a.rows / t / 1e3:.3e} Mrows/s")

    measure("prepared")
    if a.profile:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        eng.gram_pass(data, coef)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    if eng.info.cell_roles and not a.only_gram:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.refine_order(data)
        e1.record()
        torch.cuda.synchronize()
        print(f"  pgb_gram_refine: {e0.elapsed_time(e1):.3f} ms")
        measure("prepared + refined")
    if a.only_gram:
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, fn in [("gram_pass(all kernels)", lambda: eng.gram_pass(data, coef)),
                     ("stats_pass", lambda: eng.stats_pass(data, coef)),
                     ("solve", lambda: eng.solve(coef_prev=coef)),
                     ("solve+edof+cov", lambda: eng.solve(coef_prev=coef, want_edof=True, want_cov=True))]:
        fn(); fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"  {name}: {e0.elapsed_time(e1) / a.reps:.3f} ms")


if __name__ == "__main__":
    main()
