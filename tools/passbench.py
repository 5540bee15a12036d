"""Times the O(k)-per-row passes (statistics / predict) of a parity-case model straight through the C ABI
with CUDA events (no Python work inside the timed region) and reports them against the HBM roofline:
algorithmic bytes = 8 F (+ 8 for y, + 8 per written vector) per row.  Development tool.

    python tools/passbench.py --case c4_linear --rows 32000000
"""
import argparse
import contextlib
import ctypes as C
import io
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402
import helpers  # noqa: E402
from pygam_b200 import _lib  # noqa: E402

GEN = {"c2_logistic": (cases.data_c2, 10), "c3_poisson": (cases.data_c3, 24), "c4_linear": (cases.data_c4, 3),
       "c5_expectile": (cases.data_c5, 4)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="c4_linear")
    ap.add_argument("--rows", type=int, default=32_000_000)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--profile", action="store_true", help="cudaProfilerStart/Stop around one call of each pass "
                    "(ncu --profile-from-start off)")
    a = ap.parse_args()
    gen, F = GEN[a.case]
    Xs, ys = gen(20000)
    gam = helpers.build_model(cases.FIT_CASES[a.case], max_iter=2)
    with contextlib.redirect_stdout(io.StringIO()):
        gam.fit(Xs, ys)
    eng = gam._engine
    lib = _lib.load()
    n = a.rows
    reps = (n + len(ys) - 1) // len(ys)
    Xt = torch.from_numpy(np.ascontiguousarray(Xs.T)).cuda().repeat(1, reps)[:, :n].contiguous()
    y = torch.from_numpy(ys).cuda().repeat(reps)[:n].contiguous()
    eng.beta.copy_(torch.as_tensor(gam.coef_, dtype=torch.float64))
    mu = torch.empty(n, dtype=torch.float64, device="cuda")
    ws = eng._stats_ws
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)  # noqa: E731
    peak = 6544.3
    with contextlib.suppress(Exception):
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])

    def run(name, yy, mu_out, bytes_per_row):
        def call():
            _lib.check(lib.pgb_stats_pass(eng.handle, P(Xt), n, n, P(yy), C.c_void_p(0), P(eng.beta), -1.0, 0.0, 0,
                                          P(eng.stat_scalars), C.c_void_p(0), P(mu_out), P(ws), ws.numel() * 8, st), name)
        for _ in range(3):
            call()
        if a.profile:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            call()
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        gbs = n * bytes_per_row / (ms * 1e-3) / 1e9
        print(f"{a.case} n={n} {name}: {ms:.3f} ms  {n / ms / 1e3:.0f} Mrows/s  {gbs:.0f} GB/s algorithmic "
              f"= {100 * gbs / peak:.1f}% of {peak:.0f} GB/s")

    run("statistics pass (deviance sums)", y, None, 8 * F + 8)
    run("predict_mu (mu written)", None, mu, 8 * F + 8)


if __name__ == "__main__":
    main()
