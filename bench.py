#!/usr/bin/env python
"""bench.py -- rows/sec per PIRLS iteration (BASELINE.json's metric) on N B200s.

A step is ONE PIRLS iteration of BASELINE config 2 (LogisticGAM, 10 s() terms, n_splines=25,
1e7 synthetic rows per GPU): fused Gram pass over all rows (basis on the fly + link/variance +
X^T W X, X^T W z), all-reduce of the packed [G | r | scalars] buffer over the row shards, and
the m x m penalised solve.  Inputs are resident in HBM for `value`; `e2e` times the same
iteration through the host-buffer C-ABI entry point (pinned host X/y -> device inside the timed
region, coefficients read back); `e2e.fit` is a whole LogisticGAM.fit(numpy X, numpy y) from host
arrays to coef_ and statistics_.  `configs` carries the other BASELINE configurations (config 3 at
1.25e8 rows per GPU, the north star's 20 s() shape, config 4's 11^3 grid search at 5e7 rows,
config 5 at 1e8 rows, config 1 on the wage data), each with a parity check of a row prefix
against the CPU oracle.  `--impl reference` times the UNMODIFIED reference (pyGAM installed into
baseline/_ref) per PIRLS iteration on a bounded row sample, with all host cores.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 bench.py --gpus 8 --steps 10 --warmup 3
"""
import os

import sys

# The CPU arm must use every host core: torchrun exports OMP_NUM_THREADS=1 to its ranks, which made the N > 1
# reference runs of round 1 2.2x slower than the N = 1 run.  Set before numpy / BLAS load.  The ranks of a multi-GPU
# run of the B200 arm share the host instead (their numpy work is m x m bookkeeping; 8 ranks x 16 spinning BLAS
# threads on 16 cores cost the N = 8 fits of round 2 seconds of wall time).
_CORES = os.cpu_count() or 1
_REFERENCE_ARM = any(a == "reference" or a == "--impl=reference" for a in sys.argv[1:])
_THREADS = _CORES if _REFERENCE_ARM else max(1, _CORES // max(1, int(os.environ.get("WORLD_SIZE", "1"))))
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_v] = str(_THREADS)

import argparse  # noqa: E402
import contextlib  # noqa: E402
import ctypes as C  # noqa: E402
import io  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

N_ROWS_PER_GPU = 10_000_000
N_FEATURES = 10
N_SPLINES = 25
BYTES_PER_ROW = 8 * N_FEATURES + 8  # every feature column and y read once, fp64
K_NNZ = 4 * N_FEATURES + 1
FMA_PER_ROW = K_NNZ * (K_NNZ + 1) // 2 + K_NNZ  # upper triangle of the rank-1 update + rhs
FP64_FMA_PER_CLK_SM = 62.7  # measured (tools/ubench.cu, profiles/r1_ubench_fp64_dmma_smem_rmw.log)
SM_CLOCK_HZ = 1.965e9
N_SMS = 148


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=N_ROWS_PER_GPU, help="rows per GPU")
    ap.add_argument("--cpu-rows", type=int, default=100_000, help="rows of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--configs", default="auto",
                    help="other BASELINE configurations measured into the `configs` block: comma list of "
                         "c1,c3,s20,c4,c5,c2strong, 'none', or 'auto' (all at N=1; c3,c4,c2strong at N>1)")
    ap.add_argument("--config-scale", type=float, default=1.0, help="scale the rows of the `configs` block (tests)")
    return ap.parse_args()


def config(world, rows, prepare_ms=None):
    extra = {}
    if prepare_ms is not None:
        extra = {"prepare_ms_once_per_fit": prepare_ms,
                 "prepared": "the knot cell of every (row, term) and the cell order of the rows of every term pair do not "
                             "depend on the coefficients: pgb_gram_prepare builds them once per fit (like the reference "
                             "builds its model matrix once, pygam.py:741) and every timed iteration reuses them; the "
                             "streaming path of e2e sorts inside the kernel instead"}
    return {**_config(world, rows), **extra}


def _config(world, rows):
    return {"workload": "BASELINE configs[1]: LogisticGAM, 10 s() terms, n_splines=25 (m=251, k=41), "
                        f"{rows} synthetic rows per GPU, one PIRLS iteration per step",
            "rows_per_gpu": rows, "global_rows": rows * world, "n_features": N_FEATURES, "m": 251,
            "parallelism": f"row-shard x{world}, all-reduce of [G|r|scalars]" if world > 1 else "single GPU",
            "cache": "inputs (%.0f MB per GPU) larger than the 126 MB L2" % (rows * BYTES_PER_ROW / 1e6)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (pyGAM from baseline/_ref), timed per PIRLS iteration with a CallBack
# (pygam.py:838-840); the oracle port only if the reference cannot be imported
# ---------------------------------------------------------------------------------------------
def blas_threads():
    try:
        from threadpoolctl import threadpool_info

        return {f"{p.get('internal_api')}": p.get("num_threads") for p in threadpool_info()}
    except Exception:
        return {}


def import_reference():
    """pyGAM as installed by `pip install --no-index --no-deps --target baseline/_ref /root/reference`,
    else the read-only checkout (authoring container only); None when neither is there"""
    for path in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(path, "pygam")):
            for p in (os.path.join(ROOT, "baseline", "stub"), path):
                if p not in sys.path:
                    sys.path.insert(0, p)
            try:
                import pygam  # noqa: F401

                return pygam, path
            except Exception as e:  # noqa: BLE001
                sys.stderr.write(f"reference at {path} not importable: {e}\n")
    return None, None


def cpu_iterations(n_rows, steps, warmup):
    import cases

    X, y = cases.data_c2(n_rows, seed=1234)
    pygam, where = import_reference()
    if pygam is not None:
        from pygam.callbacks import CallBack

        class Stamp(CallBack):
            def __init__(self):
                super().__init__(name="stamp")

            def on_loop_start(self, gam):
                return time.perf_counter()

        terms = pygam.s(0, n_splines=N_SPLINES)
        for j in range(1, N_FEATURES):
            terms = terms + pygam.s(j, n_splines=N_SPLINES)
        # tol = 0 never stops early: exactly warmup + steps + 1 iterations, stamped at loop start
        gam = pygam.LogisticGAM(terms, max_iter=warmup + steps + 1, tol=0.0, callbacks=["deviance", "diffs", Stamp()])
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            gam.fit(X, y)
        stamps = np.asarray(gam.logs_["stamp"])
        kind = "reference"
        what = f"unmodified pyGAM {getattr(pygam, '__version__', '')} ({os.path.relpath(where, ROOT) if where.startswith(ROOT) else where})"
    else:
        from oracle import gam_oracle as orc

        case = dict(cases.FIT_CASES["c2_logistic"])
        case["model_kwargs"] = dict(max_iter=warmup + steps + 1, tol=0.0)
        t0 = time.perf_counter()
        res = orc.fit(case, X, y, timer=time.perf_counter)
        stamps = np.asarray(res["logs"]["time"])
        kind = "port"
        what = "oracle port of the reference PIRLS (the reference could not be imported)"
    setup = stamps[0] - t0
    per_iter = np.diff(stamps)[warmup:warmup + steps]
    return dict(ms_per_step=float(np.mean(per_iter) * 1e3), rows_per_s=float(n_rows / np.mean(per_iter)),
                setup_s=float(setup), iters=len(per_iter), kind=kind, what=what, blas=blas_threads())


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    r = cpu_iterations(args.cpu_rows, steps, warmup)
    sample = (f"{r['what']}: {args.cpu_rows} rows of the workload (seeded like the GPU arm), {steps} timed PIRLS iterations "
              f"after {warmup} warm-up, stamped by a CallBack at on_loop_start; model-matrix setup {r['setup_s']:.1f}s "
              f"excluded; BLAS threads {r['blas']}, os.cpu_count() {_CORES}")
    cfg = _config(1, args.cpu_rows)
    cfg["workload"] = ("BASELINE configs[1] shape: LogisticGAM, 10 s() terms, n_splines=25 (m=251, k=41), "
                       f"{args.cpu_rows} synthetic rows on the host (the reference needs 15 KB of RAM per row: the GPU arm's "
                       f"{args.rows} rows per GPU do not fit any host), one PIRLS iteration per step")
    cfg["parallelism"] = f"single process, {_CORES} BLAS threads"
    cfg["gpu_arm_rows_per_gpu"] = args.rows
    line = {"metric": "rows/sec per PIRLS iteration", "value": r["rows_per_s"], "unit": "rows/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": r["rows_per_s"], "unit": "rows/s", "cores": _CORES, "kind": r["kind"],
                             "blas_threads": r["blas"], "sample": sample},
            "e2e": {"value": r["rows_per_s"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s"


# ---- the other BASELINE configurations -----------------------------------------------------------------------------
def _terms(pg, n_s, te_pairs=(), n_splines=None):
    terms = None
    for j in range(n_s):
        t = pg.s(j) if n_splines is None else pg.s(j, n_splines=n_splines)
        terms = t if terms is None else terms + t
    for a, b in te_pairs:
        terms = terms + pg.te(a, b)
    return terms


def synth(torch, dev, name, n, seed):
    """SURVEY.md 8d generators, on the device (feature-major)"""
    gen = torch.Generator(device=dev).manual_seed(seed)
    two_pi = 2 * np.pi
    if name in ("c3", "s20"):
        F = 24 if name == "c3" else 20
        Xt = torch.rand((F, n), dtype=torch.float64, device=dev, generator=gen)
        eta = torch.full((n,), 0.5, dtype=torch.float64, device=dev)
        for j in range(20):
            eta += 0.075 * torch.sin(two_pi * Xt[j])
        if name == "c3":
            eta += 0.5 * Xt[20] * Xt[21] - 0.5 * Xt[22] * Xt[23]
        y = torch.poisson(torch.exp(eta), generator=gen)
    elif name in ("c4",):
        Xt = torch.rand((3, n), dtype=torch.float64, device=dev, generator=gen)
        y = torch.sin(two_pi * Xt[0]) + Xt[1] ** 2 + 0.5 * torch.cos(2 * two_pi * Xt[2])
        y += 0.3 * torch.randn(n, dtype=torch.float64, device=dev, generator=gen)
    elif name == "c5":
        Xt = torch.rand((4, n), dtype=torch.float64, device=dev, generator=gen)
        y = torch.sqrt(Xt[0]) + torch.sin(3 * Xt[1]) + Xt[2] ** 2 + 0.2 * torch.cos(6 * Xt[3])
        y += 0.3 * torch.randn(n, dtype=torch.float64, device=dev, generator=gen)
    else:  # c2
        Xt = torch.rand((N_FEATURES, n), dtype=torch.float64, device=dev, generator=gen)
        eta = torch.zeros(n, dtype=torch.float64, device=dev)
        for j in range(N_FEATURES):
            eta += 0.5 * torch.sin(two_pi * (j + 1) * Xt[j] / 3.0)
        y = (torch.rand(n, dtype=torch.float64, device=dev, generator=gen) < torch.sigmoid(eta)).double()
    return Xt, y


def make_model(pg, name):
    if name == "c3":
        return pg.PoissonGAM(_terms(pg, 20, [(20, 21), (22, 23)])), dict(model="PoissonGAM", terms="c3")
    if name == "s20":
        return pg.PoissonGAM(_terms(pg, 20)), dict(model="PoissonGAM", terms="s20")
    if name == "c4":
        return pg.LinearGAM(_terms(pg, 3)), dict(model="LinearGAM", terms="c4")
    if name == "c5":
        t = (pg.s(0, constraints="monotonic_inc") + pg.s(1) + pg.s(2, constraints="monotonic_inc")
             + pg.s(3, constraints="convex"))
        return pg.ExpectileGAM(t, expectile=0.9), dict(model="ExpectileGAM", terms="c5")
    if name == "c1":
        return pg.LinearGAM(pg.s(0) + pg.s(1) + pg.f(2)), dict(model="LinearGAM", terms="c1")
    return pg.LogisticGAM(_terms(pg, N_FEATURES, n_splines=N_SPLINES)), dict(model="LogisticGAM", terms="c2")


def oracle_case(name):
    import cases

    s, te = cases.s, cases.te
    if name == "c3":
        return cases.FIT_CASES["c3_poisson"]
    if name == "s20":
        return dict(model="PoissonGAM", terms=[s(j) for j in range(20)])
    if name == "c4":
        return cases.FIT_CASES["c4_linear"]
    if name == "c5":
        return cases.FIT_CASES["c5_expectile"]
    if name == "c1":
        return cases.FIT_CASES["c1_wage"]
    del te
    return cases.FIT_CASES["c2_logistic"]


WORK = {  # name: (rows per GPU, features, non-zeros per model-matrix row, parity prefix, description)
    "c3": (125_000_000, 24, 113, 10_000, "BASELINE configs[2]: PoissonGAM, 20 s() + 2 te() (m=601, k=113)"),
    "s20": (125_000_000, 20, 81, 10_000, "north star shape: PoissonGAM, 20 s() terms (m=401, k=81)"),
    "c4": (50_000_000, 3, 13, 100_000, "BASELINE configs[3]: LinearGAM gridsearch over an 11^3 lam grid on 3 s() terms (m=61)"),
    "c5": (100_000_000, 4, 17, 20_000, "BASELINE configs[4]: ExpectileGAM(0.9), monotonic / convex constrained s() terms (m=81)"),
    "c1": (3000, 3, 10, 3000, "BASELINE configs[0]: LinearGAM(s(0)+s(1)+f(2)) on the wage data (m=46)"),
}


def parity_prefix(pg, name, Xh, yh, quiet):
    """the engine and the CPU oracle on the same row prefix -- identical iteration count, GCV / UBRE
    to 1e-10 (constrained: 1e-7), edof"""
    from oracle import gam_oracle as orc

    from pygam_b200.engine import single_rank

    gam, _ = make_model(pg, name)
    t0 = time.perf_counter()
    with quiet(), single_rank():  # rank 0 alone fits the prefix: no collective inside
        gam.fit(Xh, yh)
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    with quiet(), np.errstate(all="ignore"):
        ref = orc.fit(oracle_case(name), Xh, yh)
    t_cpu = time.perf_counter() - t0
    st, rst = gam.statistics_, ref["statistics"]
    obj = "UBRE" if gam.distribution._known_scale else "GCV"
    rel = abs(st[obj] - rst[obj]) / abs(rst[obj])
    tol = 1e-7 if name == "c5" else 1e-10
    ok = (len(gam.logs_["diffs"]) == len(ref["logs"]["diffs"])) and rel <= tol and abs(st["edof"] - rst["edof"]) <= 1e-6 * rst["edof"]
    return {"rows": int(len(yh)), "ok": bool(ok), "iterations": [len(gam.logs_["diffs"]), len(ref["logs"]["diffs"])],
            obj + "_rel_diff": float(rel), "edof": [float(st["edof"]), float(rst["edof"])],
            "oracle_fit_s": t_cpu, "engine_fit_s_incl_upload": t_gpu}


def cells_kernel_names(info):
    """the kernels of a prepared cell-owned Gram pass of this model, as the launch list names them"""
    th = info.cell_threads
    pair = f"gram_cells_ring_kernel<{th}, {info.cell_ring_groups}>" if info.cell_ring_groups > 0 else f"gram_cells_sorted_kernel<{th}, 2>"
    irls = "stats_pass_kernel<1, 2, ...> (the IRLS step: omega, omega z per row)"
    if info.cell_generic_roles > 0:
        return f"{irls} + {pair} + gram_cells_generic_kernel"
    return f"{irls} + {pair}"


def run_config(name, args, torch, dist, pg, dev, rank, world, hbm_peak):
    """one entry of the `configs` block: full fit (or grid search) on device-resident rows, the steady-state
    iteration against both roofs, a parity prefix"""
    from pygam_b200 import _lib
    from pygam_b200.engine import DeviceData

    rows, F, k, n_ref, desc = WORK[name]
    n = rows if name == "c1" else max(20_000, int(rows * args.config_scale))

    def quiet():
        return contextlib.redirect_stdout(io.StringIO())

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = {"workload": desc, "rows_per_gpu": n, "global_rows": n * world}
    if name == "c1":
        w = np.load(os.path.join(ROOT, "tests", "golden", "wage_xy.npz"))
        Xh, yh = w["X"], w["y"]
        data = DeviceData.from_host(Xh, yh)
        if world > 1:
            return None
    else:
        free, _ = torch.cuda.mem_get_info()
        need = n * (8 * F + 8) * 1.3 + n * (9 * F + 16 + 2.2 * (F * (F + 1) // 2 if name != "c3" else 260))
        if need > 0.92 * free:
            n = int(n * 0.92 * free / need) // 1_000_000 * 1_000_000
            out["rows_per_gpu"], out["global_rows"] = n, n * world
            out["note"] = "rows reduced to fit the free HBM"
        Xt, y = synth(torch, dev, name, n, 4321 + 17 * rank)
        data = DeviceData(Xt, y)
        Xh = Xt[:, :n_ref].T.contiguous().cpu().numpy()
        yh = y[:n_ref].cpu().numpy()
    bytes_row = 8 * F + 8
    fma_row = k * (k + 1) // 2 + k
    gam, _ = make_model(pg, name)
    sync()
    t0 = time.perf_counter()
    if name == "c4":
        with quiet():
            gam.gridsearch(data, lam=[np.logspace(-3, 3, 11)] * 3)
        sync()
        out["gridsearch_wall_s"] = time.perf_counter() - t0
        out["candidates"] = 1331
        out["how"] = ("one Gram pass over the rows, one O(k) pass for the moments of y, 1331 m x m solves as one batch "
                      f"split over {world} rank(s), scores from quadratic forms (GAM._gridsearch_linear_batched)")
        out["best_lam"] = [float(v) for v in np.ravel(gam.lam)]
        out["best_GCV"] = float(gam.statistics_["GCV"])
    else:
        with quiet():
            gam.fit(data)
        sync()
        out["fit_wall_s"] = time.perf_counter() - t0
        out["iterations"] = len(gam.logs_["diffs"])
        out["converged"] = bool(gam.logs_["diffs"][-1] < gam.tol)
        obj = "UBRE" if gam.distribution._known_scale else "GCV"
        out[obj] = float(gam.statistics_[obj])
        out["edof"] = float(gam.statistics_["edof"])
    # steady-state iteration: Gram pass (+ all-reduce) + solve at the fitted coefficients, CUDA events
    eng = gam._engine
    coef = np.array(gam.coef_)
    reps = 3
    # the optional second preparation step (the engine runs it by itself after 50 passes over a prepared shard): here,
    # outside the timed iterations, and reported
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.refine_order(data)
    e1.record()
    sync()
    out["refine_ms_once_per_long_fit"] = e0.elapsed_time(e1)
    for _ in range(2):
        eng.gram_pass(data, coef)
        eng.solve(coef_prev=coef)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for _ in range(reps):
        eng.gram_pass(data, coef)
        eng.solve(coef_prev=coef)
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_iter = float(ms.item())
    _lib.check(eng.lib.pgb_gram_timing(eng.handle, 1), "timing")
    eng.gram_pass(data, coef)
    f, r, t = C.c_float(), C.c_float(), C.c_float()
    _lib.check(eng.lib.pgb_gram_last_ms(eng.handle, C.byref(r), C.byref(f), C.byref(t)), "ms")
    _lib.check(eng.lib.pgb_gram_timing(eng.handle, 0), "timing")
    gram_ms = t.value
    out.update({
        "ms_per_iteration": ms_iter, "rows_per_s": n * world / (ms_iter * 1e-3), "gram_pass_ms": gram_ms,
        "algorithmic_bytes_per_row": bytes_row, "algorithmic_fma_per_row": fma_row,
        "hbm_frac": n * bytes_row / (gram_ms * 1e-3) / 1e9 / hbm_peak,
        "fp64_fma_per_clk_per_sm": n * fma_row / (gram_ms * 1e-3) / (N_SMS * SM_CLOCK_HZ),
        "fp64_frac": n * fma_row / (gram_ms * 1e-3) / (N_SMS * SM_CLOCK_HZ) / FP64_FMA_PER_CLK_SM,
        "gram_path": cells_kernel_names(eng.info) if eng.info.cell_roles > 0 else "general (shared-memory accumulation)"})
    # the O(k) passes of this shape against the HBM roof
    if name != "c1":
        for _ in range(2):
            eng.stats_pass(data, coef)
        sync()
        e0.record()
        for _ in range(5):
            eng.stats_pass(data, coef, reduce=False)
        e1.record()
        sync()
        st_ms = e0.elapsed_time(e1) / 5
        out["statistics_pass_ms"] = st_ms
        out["statistics_pass_hbm_frac"] = n * bytes_row / (st_ms * 1e-3) / 1e9 / hbm_peak
    del data
    if name != "c1":
        del Xt, y
    gam._engine = None
    eng = None
    torch.cuda.empty_cache()
    if rank == 0:
        try:
            out["parity_prefix"] = parity_prefix(pg, name, Xh, yh, quiet)
        except Exception as e:  # noqa: BLE001
            out["parity_prefix"] = {"ok": False, "error": repr(e)[:200]}
    sync()
    return out


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import pygam_b200 as pg
    from pygam_b200 import _lib
    from pygam_b200.engine import DeviceData

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime

        # a rank that fails must not leave the others waiting for ten minutes
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    n = args.rows
    lib = _lib.load()
    hbm_peak, peak_source = hbm_peak_gbs()

    # synthetic shard, generated on the device
    Xt, y = synth(torch, dev, "c2", n, 1234 + rank)
    data = DeviceData(Xt, y)
    gam, _ = make_model(pg, "c2")
    gam.max_iter = 1
    # compile the model (edge knots from the global column extrema) and get the initial estimate
    # through the public API: one PIRLS iteration
    with contextlib.redirect_stdout(io.StringIO()):
        gam.fit(data)
    eng = gam._engine
    m = eng.m
    coef = np.array(gam.coef_)
    S_P = np.eye(m) * np.sqrt(np.finfo(np.float64).eps) + gam._P()
    eng.set_penalty(S_P)

    def step(c):
        eng.gram_pass(data, c)  # fused Gram pass + all-reduce
        sol = eng.solve(coef_prev=c)  # m x m solve, coefficients read back
        return sol["coef"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # what does not depend on the coefficients (knot cells of the rows and their order) is prepared once per
    # fit by the engine; time that one-time step on its own so that the steady-state iteration can be read
    # against it (the metric is the median iteration 2..last of a fit)
    prepare_ms = refine_ms = None
    if eng.info.cell_roles > 0:
        ws = eng._ensure_gram_ws(n)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        p0.record()
        _lib.check(lib.pgb_gram_prepare(eng.handle, C.c_void_p(Xt.data_ptr()), n, data.ld, C.c_void_p(ws.data_ptr()),
                                        ws.numel() * 8, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pgb_gram_prepare")
        p1.record()
        torch.cuda.synchronize()
        prepare_ms = p0.elapsed_time(p1)
        # second, optional step of the preparation (pgb_gram_refine: bank-aware order of the rows inside a cell); the
        # engine runs it by itself once a prepared shard has been passed over 50 times, here it is run (and timed) up
        # front so that every timed step is a steady-state iteration
        p0.record()
        eng.refine_order(data)
        p1.record()
        torch.cuda.synchronize()
        refine_ms = p0.elapsed_time(p1)
    for _ in range(max(args.warmup, 3)):
        coef = step(coef)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    launches0 = lib.pgb_launch_count()
    _lib.check(lib.pgb_gram_timing(eng.handle, 1), "pgb_gram_timing")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    accum_ms, rec_ms, pass_ms = [], [], []
    e0.record()
    for _ in range(args.steps):
        coef = step(coef)
        f, r, t = C.c_float(), C.c_float(), C.c_float()
        _lib.check(lib.pgb_gram_last_ms(eng.handle, C.byref(r), C.byref(f), C.byref(t)), "pgb_gram_last_ms")
        accum_ms.append(f.value)
        rec_ms.append(r.value)
        pass_ms.append(t.value)
    e1.record()
    barrier()
    launches = lib.pgb_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    _lib.check(lib.pgb_gram_timing(eng.handle, 0), "pgb_gram_timing")
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    acc = torch.tensor([float(np.mean(accum_ms)), float(np.mean(rec_ms)), float(np.mean(pass_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(acc, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    accum, records, gram_pass_ms = float(acc[0].item()), float(acc[1].item()), float(acc[2].item())
    value = n * world / (ms_per_step * 1e-3)

    # ---- the O(k)-per-row passes (statistics sums, predict_mu): the HBM-bound kernels of the path ----
    def time_pass(fn, reps=10):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    eng.beta.copy_(torch.as_tensor(coef, dtype=torch.float64))
    mu_buf = torch.empty(n, dtype=torch.float64, device=dev)
    st_ws = eng._stats_ws
    cur = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _stats_call(yy, mu_out):
        _lib.check(lib.pgb_stats_pass(eng.handle, C.c_void_p(Xt.data_ptr()), n, data.ld, C.c_void_p(yy.data_ptr() if yy is not None else 0),
                                      C.c_void_p(0), C.c_void_p(eng.beta.data_ptr()), -1.0, 0.0, 0,
                                      C.c_void_p(eng.stat_scalars.data_ptr()), C.c_void_p(0),
                                      C.c_void_p(mu_out.data_ptr() if mu_out is not None else 0), C.c_void_p(st_ws.data_ptr()),
                                      st_ws.numel() * 8, cur), "pgb_stats_pass")

    stats_ms = time_pass(lambda: _stats_call(y, None))
    predict_ms = time_pass(lambda: _stats_call(None, mu_buf))
    del mu_buf

    # ---- end to end: host buffers in, coefficients out, every step ----
    e2e = None
    if not args.no_e2e:
        Xh = torch.empty((N_FEATURES, n), dtype=torch.float64).pin_memory()
        yh = torch.empty(n, dtype=torch.float64).pin_memory()
        Xh.copy_(Xt)
        yh.copy_(y)
        out_h = torch.empty(m * m + m + _lib.GS_COUNT, dtype=torch.float64).pin_memory()

        def e2e_step(c):
            cc = np.ascontiguousarray(c)
            _lib.check(lib.pgb_gram_pass_host(eng.handle, _lib.MODE_IRLS, C.c_void_p(Xh.data_ptr()), n, n,
                                              C.c_void_p(yh.data_ptr()), C.c_void_p(0),
                                              cc.ctypes.data_as(C.c_void_p), C.c_void_p(out_h.data_ptr()), 1 << 20),
                       "pgb_gram_pass_host")
            eng.out.copy_(out_h, non_blocking=True)
            if world > 1:
                dist.all_reduce(eng.out)
            return eng.solve(coef_prev=c)["coef"]

        ce = coef
        for _ in range(2):
            ce = e2e_step(ce)
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(3, min(args.steps, 5))
        for _ in range(e2e_steps):
            ce = e2e_step(ce)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_ms = float(dt.item()) * 1e3 / e2e_steps
        e2e = {"value": n * world / (e2e_ms * 1e-3), "unit": "rows/s", "ms_per_step": e2e_ms, "steps": e2e_steps,
               "h2d_bytes_per_step": int(n * BYTES_PER_ROW + m * 8), "d2h_bytes_per_step": int(out_h.numel() * 8 + (m + 8) * 8),
               "api": "pgb_gram_pass_host (pinned host X, y -> chunked H2D on 2 streams overlapped with the Gram "
                      "kernels) + all-reduce + pgb_solve, coefficients read back; every step re-uploads the rows, which "
                      "no fit does: e2e.fit below is the call a user makes",
               "note_scaling": "N ranks stream N x 880 MB per step through one host: the host memory / PCIe root bandwidth "
                               "(not NCCL, not the kernels) bounds this leg at N >= 4 (DESIGN.md 7)"}
        # the call a user makes: numpy arrays in, fitted model out (upload once, prepare, k iterations, statistics)
        if world == 1:
            Xn = np.ascontiguousarray(Xh.numpy().T)  # (n, F) row-major, like any user's array
            yn = yh.numpy().copy()
            del Xh
            with contextlib.redirect_stdout(io.StringIO()):
                make_model(pg, "c2")[0].fit(Xn[:200_000], yn[:200_000])  # warm-up of the code paths (small buffers only)
                walls = []
                for _ in range(3):
                    fit_gam = None  # the buffers of the previous fit go back to torch's caching allocator
                    fit_gam, _ = make_model(pg, "c2")
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    fit_gam.fit(Xn, yn)
                    torch.cuda.synchronize()
                    walls.append(time.perf_counter() - t0)
            fit_s = min(walls[1:])
            iters = len(fit_gam.logs_["diffs"])
            e2e["fit"] = {"api": "LogisticGAM(10 x s(n_splines=25)).fit(X_numpy, y_numpy): host arrays -> coef_, statistics_",
                          "rows": n, "wall_s": fit_s, "first_fit_wall_s": walls[0],
                          "note": "wall_s: best of the second and third fit of this size; first_fit_wall_s: the first one in "
                                  "this process (where torch's caching allocator has no blocks of these sizes yet it also pays "
                                  "the cudaMalloc of ~4 GB of buffers: tools/fitstages.py)",
                          "iterations": iters, "rows_per_s_per_iteration": n * iters / fit_s,
                          "h2d_bytes": int(n * BYTES_PER_ROW), "UBRE": float(fit_gam.statistics_["UBRE"]),
                          "edof": float(fit_gam.statistics_["edof"])}
            del Xn, yn, fit_gam
        else:
            del Xh
        del yh

    del data, Xt, y
    pass_kernels = cells_kernel_names(eng.info) if eng.info.cell_roles > 0 else "gram_records_kernel + gram_accum_kernel"
    gam._engine = None
    eng = None
    torch.cuda.empty_cache()
    # ---- the other BASELINE configurations ----
    which = args.configs
    if which == "auto":
        which = "c1,c3,s20,c4,c5" if world == 1 else "c3,c4,c2strong"
    configs = {}
    for name in [w for w in which.split(",") if w and w != "none"]:
        try:
            if name == "c2strong":
                configs[name] = strong_point(args, torch, dist, pg, dev, rank, world)
            else:
                res = run_config(name, args, torch, dist, pg, dev, rank, world, hbm_peak)
                if res is not None:
                    configs[name] = res
        except Exception as e:  # noqa: BLE001
            configs[name] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    gram_ms = gram_pass_ms  # every kernel of the Gram pass (prepass, cell sums, reductions)
    achieved = n * BYTES_PER_ROW / (gram_ms * 1e-3) / 1e9
    fma_per_clk_sm = n * FMA_PER_ROW / (accum * 1e-3) / (N_SMS * SM_CLOCK_HZ)
    traffic = None
    traffic_src = None
    for fn in ("r2_ncu_gram_cells_traffic.json", "r1_ncu_gram_cells_traffic.json"):
        try:  # dram bytes of the two kernels per launch from the committed ncu --set full capture, scaled to n rows
            prof = json.load(open(os.path.join(ROOT, "profiles", fn)))
            traffic, traffic_src = prof["dram_bytes_per_row"] * n, "profiles/" + fn
            break
        except Exception:
            pass
    kernels = pass_kernels + " (the Gram pass of one PIRLS iteration of a prepared shard)"
    second = {"name": "fp64_pipe", "algorithmic_fma_per_row": FMA_PER_ROW, "achieved_fma_per_clk_per_sm": fma_per_clk_sm,
              "measured_peak_fma_per_clk_per_sm": FP64_FMA_PER_CLK_SM, "frac": fma_per_clk_sm / FP64_FMA_PER_CLK_SM,
              "note": "spline-pair kernel: a thread owns a knot cell of a term pair and sums the moments of the "
                      "cell's rows in registers (row order by cell prepared once per fit); the useful work is 902 fp64 FMA "
                      "per row at 88 algorithmic bytes per row = 20 flop/B, so the fp64 pipe (62.7 FMA/clk/SM measured, "
                      "tools/ubench.cu), not HBM, bounds the pass (DESIGN.md section 4)"}
    roofline = {"kernel": kernels, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "peak_source": peak_source,
                "traffic": traffic, "traffic_source": traffic_src, "kernel_ms": gram_ms, "prepass_kernel_ms": records,
                "accum_kernel_ms": accum, "kernel_share_of_step": gram_ms / ms_per_step,
                "algorithmic_bytes_per_row": BYTES_PER_ROW, "governing": second,
                "ok_passes": {"note": "the O(k)-per-row passes of the path, same shard and coefficients, CUDA events over 10 "
                                      "launches each: statistics sums (stats_pass_kernel, reads 8F+8 B per row) and predict_mu "
                                      "(reads 8F, writes 8 B per row)",
                              "statistics_pass_ms": stats_ms,
                              "statistics_pass_gbs": n * BYTES_PER_ROW / (stats_ms * 1e-3) / 1e9,
                              "statistics_pass_frac": n * BYTES_PER_ROW / (stats_ms * 1e-3) / 1e9 / hbm_peak,
                              "predict_mu_ms": predict_ms,
                              "predict_mu_gbs": n * BYTES_PER_ROW / (predict_ms * 1e-3) / 1e9,
                              "predict_mu_frac": n * BYTES_PER_ROW / (predict_ms * 1e-3) / 1e9 / hbm_peak}}
    line = {"metric": "rows/sec per PIRLS iteration", "value": value, "unit": "rows/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {**config(world, n, prepare_ms), "refine_ms_once_per_long_fit": refine_ms},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline}
    if e2e is not None:
        line["e2e"] = e2e
    if configs:
        line["configs"] = configs
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_iterations(args.cpu_rows, 3, 1)
        line["cpu_baseline"] = {"value": r["rows_per_s"], "unit": "rows/s", "cores": _CORES, "kind": r["kind"],
                                "blas_threads": r["blas"],
                                "sample": f"{r['what']} on {args.cpu_rows} rows of the same synthetic workload, 3 timed "
                                          f"iterations after 1 warm-up (CallBack stamps), {r['ms_per_step']:.0f} ms per "
                                          f"iteration, setup {r['setup_s']:.1f}s excluded"}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def strong_point(args, torch, dist, pg, dev, rank, world):
    """config 2 with a FIXED total of 1e7 rows over the ranks: what the fixed cost per iteration (solve front, all-reduce,
    host synchronisation) does to a strong-scaled run"""
    from pygam_b200.engine import DeviceData

    total = int(10_000_000 * args.config_scale)
    n = total // world
    Xt, y = synth(torch, dev, "c2", n, 999 + rank)
    data = DeviceData(Xt, y)
    gam, _ = make_model(pg, "c2")
    gam.max_iter = 2
    with contextlib.redirect_stdout(io.StringIO()):
        gam.fit(data)
    eng, coef = gam._engine, np.array(gam.coef_)
    for _ in range(3):
        eng.gram_pass(data, coef)
        eng.solve(coef_prev=coef)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        eng.gram_pass(data, coef)
        eng.solve(coef_prev=coef)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 10], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    del data, Xt, y
    gam._engine = None
    torch.cuda.empty_cache()
    return {"workload": "config 2, strong scaling: 1e7 rows in total", "global_rows": n * world, "rows_per_gpu": n,
            "ms_per_iteration": float(ms.item()), "rows_per_s": n * world / (float(ms.item()) * 1e-3)}


def _emit(line):
    """the ONE JSON line goes to the process's original stdout"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# libraries loaded later (NCCL prints its version banner to stdout) must not pollute the one-line
# contract: fd 1 is pointed at stderr for the whole run, the result line is written to the saved fd
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
