"""Row sharding over 2 processes (gloo, CPU): every rank holds a contiguous row shard, the packed
[G | r | scalars] buffer, the statistics sums, the column extrema and the factor levels are
all-reduced, and every rank ends with the same model as a single-process fit on all rows.
The per-row arithmetic is the oracle-backed test engine; what is exercised here is the product's
sharding logic (engine.all_reduce_, GAM._fit_data, column_stats, count_levels)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out_dir):
    for p in (os.path.dirname(HERE), HERE, os.path.join(HERE, "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import cases
    import helpers
    from oracle_backend import OracleBackend
    from pygam_b200 import engine as pg_engine

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pg_engine.swap_backend_for_tests(OracleBackend())
    case = cases.FIT_CASES[name]
    g = helpers.load("fit_" + name)
    n = len(g["y"])
    lo, hi = n * rank // world, n * (rank + 1) // world
    shard = {k: (v[lo:hi] if k in ("X", "y", "weights", "exposure") else v) for k, v in g.items()}
    gam = helpers.fit_case(case, shard)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), coef=gam.coef_, edof=gam.statistics_["edof"],
             dev=gam.statistics_["deviance"], n_iter=len(gam.logs_["diffs"]), n=gam.statistics_["n_samples"],
             logdev=np.array(gam.logs_["deviance"]))
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["c1_wage", "lam_large", "poisson_exposure"])
def test_two_rank_row_sharding_matches_single_process(name, tmp_path):
    import cases
    import helpers
    from oracle_backend import OracleBackend
    from pygam_b200 import engine as pg_engine

    mp.spawn(_worker, args=(2, _free_port(), name, str(tmp_path)), nprocs=2, join=True)
    r0 = dict(np.load(tmp_path / "rank0.npz"))
    r1 = dict(np.load(tmp_path / "rank1.npz"))
    old = pg_engine.swap_backend_for_tests(OracleBackend())
    try:
        g = helpers.load("fit_" + name)
        single = helpers.fit_case(cases.FIT_CASES[name], g)
    finally:
        pg_engine.swap_backend_for_tests(old)
    # both ranks hold the same model (the solve is replicated and deterministic)
    np.testing.assert_array_equal(r0["coef"], r1["coef"])
    assert int(r0["n"]) == len(g["y"]) == int(r1["n"])
    assert int(r0["n_iter"]) == len(single.logs_["diffs"]) == int(g["n_iter"])
    Pz = helpers.identifiable_projector(single)
    assert np.linalg.norm(Pz @ (r0["coef"] - single.coef_)) <= 1e-9 * np.linalg.norm(single.coef_)
    assert abs(float(r0["edof"]) - single.statistics_["edof"]) <= 1e-9 * single.statistics_["edof"]
    assert abs(float(r0["dev"]) - single.statistics_["deviance"]) <= 1e-10 * abs(single.statistics_["deviance"])
    np.testing.assert_allclose(r0["logdev"], single.logs_["deviance"], rtol=1e-10)
    # and the sharded fit still matches the unmodified reference
    helpers_tol = 1e-8
    assert np.linalg.norm(Pz @ (r0["coef"] - g["coef"])) <= helpers_tol * np.linalg.norm(g["coef"])
