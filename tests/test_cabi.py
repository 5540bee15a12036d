"""The C-ABI library loads and exports every symbol include/pgb.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pgb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pgb_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for name in ("pgb_model_create", "pgb_gram_pass", "pgb_stats_pass", "pgb_solve", "pgb_solve_prepare", "pgb_gram_pass_host",
                 "pgb_col_stats", "pgb_model_matrix", "pgb_chol_check"):
        assert name in syms


def test_library_exports_every_declared_symbol():
    from pygam_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge

        ge.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/pgb.h but not exported by libpgb.so"
    # and the ctypes binding covers the same set
    assert sorted(_lib.SYMBOLS) == declared_symbols()
    assert _lib.load().pgb_version() >= 100


def test_struct_layout_matches_header():
    from pygam_b200 import _lib

    # pgb_term: 2 + 4*4 int32, 8 doubles, 4 int32 -> 72 + 64 + 16 = 152 bytes
    assert ctypes.sizeof(_lib.PgbTerm) == 152
    assert ctypes.sizeof(_lib.PgbModelInfo) == 64  # 6 int32, 2 int64, 6 int32


def test_no_cpu_fallback(monkeypatch):
    """without the shared library the product refuses to run"""
    from pygam_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libpgb.so")
    with pytest.raises(ImportError):
        _lib.load()


@pytest.mark.gpu
@pytest.mark.parametrize("n,F,threads", [(1, 1, 0), (7, 3, 1), (100_003, 10, 0), (52_430, 10, 3), (600_001, 1, 0), (2_000, 700, 2),
                                         (1_300_000, 24, 0)])
def test_upload_rows_is_a_bit_exact_transpose(n, F, threads):
    """pgb_upload_rows: (n, F) row-major host rows -> feature-major device columns, bit for bit (blocks that end inside
    the array, one block, more workers than blocks, a vector); DeviceData.from_host goes through it"""
    import numpy as np
    import torch

    from pygam_b200 import _lib, engine

    lib = _lib.load()
    rng = np.random.default_rng(n + F)
    X = rng.standard_normal((n, F))
    X[0, 0] = np.nan  # payload bits travel unchanged
    ld = n + 5
    dst = torch.full((F, ld), -1.0, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    _lib.check(lib.pgb_upload_rows(dst.data_ptr(), ld, X.ctypes.data, n, F, threads), "pgb_upload_rows")
    got = dst.cpu().numpy()
    assert np.array_equal(got[:, :n].view(np.int64), np.ascontiguousarray(X.T).view(np.int64))
    assert np.all(got[:, n:] == -1.0)
    X[0, 0] = 0.5
    y = rng.standard_normal(n)
    data = engine.DeviceData.from_host(X, y)
    assert np.array_equal(data.Xt.cpu().numpy(), X.T) and np.array_equal(data.y.cpu().numpy(), y)
    assert lib.pgb_upload_rows(dst.data_ptr(), n - 1, X.ctypes.data, n, F, 0) == _lib.EINVAL  # ld < n
