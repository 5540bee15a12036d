#!/usr/bin/env python
"""bench.py -- rows/sec per PIRLS iteration (BASELINE.json's metric) on N B200s.

A step is ONE PIRLS iteration of BASELINE config 2 (LogisticGAM, 10 s() terms, n_splines=25,
1e7 synthetic rows per GPU): fused Gram pass over all rows (basis on the fly + link/variance +
X^T W X, X^T W z), all-reduce of the packed [G | r | scalars] buffer over the row shards, and
the m x m penalised solve.  Inputs are resident in HBM for `value`; `e2e` times the same
iteration through the host-buffer C-ABI entry point (pinned host X/y -> device inside the timed
region, coefficients read back).  `--impl reference` times the CPU port of the reference
algorithm (oracle/gam_oracle.py: sparse model matrix, dense QR, SVD) on a bounded row sample.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 bench.py --gpus 8 --steps 10 --warmup 3
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

N_ROWS_PER_GPU = 10_000_000
N_FEATURES = 10
N_SPLINES = 25
BYTES_PER_ROW = 8 * N_FEATURES + 8  # every feature column and y read once, fp64
K_NNZ = 4 * N_FEATURES + 1
FMA_PER_ROW = K_NNZ * (K_NNZ + 1) // 2 + K_NNZ  # upper triangle of the rank-1 update + rhs


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
# CPU arm: the oracle port of the reference algorithm, timed per PIRLS iteration
# ---------------------------------------------------------------------------------------------
def cpu_iterations(n_rows, steps, warmup):
    import cases
    from oracle import gam_oracle as orc

    X, y = cases.data_c2(n_rows, seed=1234)
    case = dict(cases.FIT_CASES["c2_logistic"])
    # tol = 0 never stops early: exactly warmup + steps + 1 iterations, stamped at loop start
    case["model_kwargs"] = dict(max_iter=warmup + steps + 1, tol=0.0)
    t0 = time.perf_counter()
    res = orc.fit(case, X, y, timer=time.perf_counter)
    stamps = np.asarray(res["logs"]["time"])
    setup = stamps[0] - t0
    per_iter = np.diff(stamps)[warmup:warmup + steps]
    return dict(ms_per_step=float(np.mean(per_iter) * 1e3), rows_per_s=float(n_rows / np.mean(per_iter)),
                setup_s=float(setup), iters=len(per_iter))


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    r = cpu_iterations(args.cpu_rows, steps, warmup)
    cores = os.cpu_count()
    sample = (f"{args.cpu_rows} of the workload's rows (seeded like the GPU arm), {steps} timed PIRLS iterations "
              f"after {warmup} warm-up; model-matrix setup {r['setup_s']:.1f}s excluded")
    line = {"metric": "rows/sec per PIRLS iteration", "value": r["rows_per_s"], "unit": "rows/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config(args.gpus, args.rows),
            "cpu_baseline": {"value": r["rows_per_s"], "unit": "rows/s", "cores": cores, "kind": "port",
                             "sample": sample},
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


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import pygam_b200 as pg
    from pygam_b200 import _lib
    from pygam_b200.engine import DeviceData

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.rows
    lib = _lib.load()

    # synthetic shard, generated on the device
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    Xt = torch.rand((N_FEATURES, n), dtype=torch.float64, device=dev, generator=gen)
    eta = torch.zeros(n, dtype=torch.float64, device=dev)
    for j in range(N_FEATURES):
        eta += 0.5 * torch.sin(2 * np.pi * (j + 1) * Xt[j] / 3.0)
    y = (torch.rand(n, dtype=torch.float64, device=dev, generator=gen) < torch.sigmoid(eta)).double()
    del eta
    data = DeviceData(Xt, y)

    terms = None
    for j in range(N_FEATURES):
        t = pg.s(j, n_splines=N_SPLINES)
        terms = t if terms is None else terms + t
    gam = pg.LogisticGAM(terms, max_iter=1)
    # compile the model (edge knots from the global column extrema) and get the initial estimate
    # through the public API: one PIRLS iteration
    import contextlib
    import io

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
    prepare_ms = None
    if eng.info.cell_roles > 0:
        ws = eng._ensure_gram_ws(n)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        p0.record()
        _lib.check(lib.pgb_gram_prepare(eng.handle, C.c_void_p(Xt.data_ptr()), n, n, C.c_void_p(ws.data_ptr()),
                                        ws.numel() * 8, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pgb_gram_prepare")
        p1.record()
        torch.cuda.synchronize()
        prepare_ms = p0.elapsed_time(p1)
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
        _lib.check(lib.pgb_stats_pass(eng.handle, C.c_void_p(Xt.data_ptr()), n, n, C.c_void_p(yy.data_ptr() if yy is not None else 0),
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
                      "kernels) + all-reduce + pgb_solve, coefficients read back"}
        del Xh, yh

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    gram_ms = gram_pass_ms  # every kernel of the Gram pass (prepass/records, accumulation, reductions)
    achieved = n * BYTES_PER_ROW / (gram_ms * 1e-3) / 1e9
    fma_per_clk_sm = n * FMA_PER_ROW / (accum * 1e-3) / (148 * 1.965e9)
    cells = eng.info.cell_roles > 0
    traffic = None
    try:  # dram bytes of the dominant kernel per launch from the committed ncu --set full capture, scaled to n rows
        prof = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_gram_cells_traffic.json")))
        traffic = prof["dram_bytes_per_row"] * n
    except Exception:
        pass
    if cells:
        kernels = "gram_prepass_kernel + gram_cells_kernel (the Gram pass of one PIRLS iteration)"
        second = {"name": "fp64_pipe", "algorithmic_fma_per_row": FMA_PER_ROW, "achieved_fma_per_clk_per_sm": fma_per_clk_sm,
                  "measured_peak_fma_per_clk_per_sm": 62.7, "frac": fma_per_clk_sm / 62.7,
                  "note": "gram_cells_kernel: a thread owns a knot cell of a term pair and sums the 4 x 4 outer products of "
                          "the cell's rows in registers (row order by cell prepared once per fit; the streaming path counting-sorts in shared memory); the useful work is "
                          "902 fp64 FMA per row at 88 algorithmic bytes per row = 20 flop/B, so the fp64 pipe (62.7 "
                          "FMA/clk/SM measured, tools/ubench.cu), not HBM, bounds the pass (DESIGN.md section 4)"}
    else:
        kernels = "gram_records_kernel + gram_accum_kernel (the Gram pass of one PIRLS iteration)"
        second = {"name": "shared_memory_rmw", "algorithmic_fma_per_row": FMA_PER_ROW,
                  "achieved_fma_per_clk_per_sm": fma_per_clk_sm, "measured_peak_fma_per_clk_per_sm": 7.7,
                  "frac": fma_per_clk_sm / 7.7,
                  "note": "general path: products scattered into G held in shared memory, 16 B of read-modify-write per FMA"}
    roofline = {"kernel": kernels, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json (measured copy bandwidth)" if peaks else "fallback 6650 GB/s",
                "traffic": traffic, "kernel_ms": gram_ms, "prepass_kernel_ms": records, "accum_kernel_ms": accum,
                "kernel_share_of_step": gram_ms / ms_per_step, "algorithmic_bytes_per_row": BYTES_PER_ROW,
                "governing": second,
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
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config(world, n, prepare_ms),
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline}
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_iterations(args.cpu_rows, 3, 1)
        line["cpu_baseline"] = {"value": r["rows_per_s"], "unit": "rows/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"oracle port of the reference PIRLS (sparse model matrix + dense QR + SVD) "
                                          f"on {args.cpu_rows} rows of the same synthetic workload, 3 timed iterations "
                                          f"after 1 warm-up, {r['ms_per_step']:.0f} ms per iteration"}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


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
