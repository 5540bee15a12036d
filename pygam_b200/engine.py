"""Device engine: owns the row shard in HBM and drives the C-ABI kernels of ``libpgb.so``.

One ``Engine`` = one compiled model descriptor (``pgb_model``) + the workspaces of the fused
passes on one GPU.  PyTorch is used only for device memory, streams and ``torch.distributed``
(NCCL) plumbing; every numerical step on the n rows is a hand-written CUDA kernel reached
through ``ctypes``.  With ``world_size > 1`` each rank holds a contiguous row shard and the
packed ``[G | r | scalars]`` buffer is all-reduced once per PIRLS iteration.
"""

import ctypes as C

import numpy as np
import torch

from . import _lib
from .terms import FactorTerm, Intercept, LinearTerm, SplineTerm, TensorTerm


def _ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class DeviceData:
    """A row shard resident in HBM, FEATURE-MAJOR: ``Xt`` is a contiguous (F, n) float64 CUDA
    tensor (feature f of row i at ``Xt[f, i]``), ``y`` (n,) float64, ``w`` (n,) float32 or None
    (pyGAM casts user weights to float32, pygam.py:889)."""

    def __init__(self, Xt, y=None, w=None):
        if not (isinstance(Xt, torch.Tensor) and Xt.dtype == torch.float64 and Xt.dim() == 2):
            raise ValueError("Xt must be a 2-d float64 tensor of shape (n_features, n_rows)")
        # rows of Xt (feature columns) must be unit-stride; the leading dimension may be padded (from_host pads it to
        # a multiple of 16 rows so that every column starts 16-byte aligned: the TMA bulk copies of the O(k) passes)
        ok = Xt.shape[1] <= 1 or (Xt.stride(1) == 1 and (Xt.shape[0] <= 1 or Xt.stride(0) >= Xt.shape[1]))
        self.Xt = Xt if ok else Xt.contiguous()
        self.n_features, self.n = self.Xt.shape
        self.ld = self.Xt.stride(0) if self.n_features > 1 else self.n
        self.y = None if y is None else y.to(device=Xt.device, dtype=torch.float64).contiguous()
        self.w = None if w is None else w.to(device=Xt.device, dtype=torch.float32).contiguous()
        if self.y is not None and self.y.numel() != self.n:
            raise ValueError(f"Inconsistent input and output data shapes. found X: {(self.n, self.n_features)} "
                             f"and y: {tuple(self.y.shape)}")
        if self.w is not None and self.w.numel() != self.n:
            raise ValueError("weights should have shape {}, but found {}".format((self.n,), tuple(self.w.shape)))
        self._memo = {}  # column statistics / level counts of this shard, computed once

    @property
    def device(self):
        return self.Xt.device

    @property
    def shape(self):
        return (self.n, self.n_features)

    @classmethod
    def from_host(cls, X, y=None, w=None, device=None):
        """numpy (n, F) row-major -> feature-major device tensors"""
        device = get_backend().default_device() if device is None else torch.device(device)
        X = np.asarray(X)
        if X.ndim == 1:
            X = X[:, None]
        n, F = X.shape
        ld = (n + 15) // 16 * 16
        if device.type != "cuda":
            # a host-side container for the CPU test engine (tests/oracle_backend.py); the product's passes refuse it
            # (_cuda_column_stats, vector_range, the engine's kernels): there is no CPU path
            buf = torch.zeros((F, ld), dtype=torch.float64, device=device)
            Xt = buf[:, :n]
            Xt.copy_(torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).T)
            yt = None if y is None else torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64).ravel()).to(device)
            wt = None if w is None else torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32).ravel()).to(device)
            return cls(Xt, yt, wt)
        # pgb_upload_rows: worker threads transpose blocks of rows into pinned staging buffers and send each block with
        # one strided copy -- a pageable cudaMemcpy of the rows plus a device-side transpose cost more than the PIRLS
        # iterations of a 1e7-row fit
        lib = _lib.load()
        Xc = np.ascontiguousarray(X, dtype=np.float64)
        with torch.cuda.device(device):
            buf = torch.empty((F, ld), dtype=torch.float64, device=device)
            if ld > n:
                buf[:, n:].zero_()
            yt = None
            if y is not None:
                yc = np.ascontiguousarray(y, dtype=np.float64).ravel()
                yt = torch.empty(len(yc), dtype=torch.float64, device=device)
            torch.cuda.current_stream().synchronize()  # the workers copy on streams of their own
            if n:
                _lib.check(lib.pgb_upload_rows(buf.data_ptr(), ld, Xc.ctypes.data, n, F, 0), "pgb_upload_rows")
            if yt is not None and len(yc):
                _lib.check(lib.pgb_upload_rows(yt.data_ptr(), len(yc), yc.ctypes.data, len(yc), 1, 0), "pgb_upload_rows")
        Xt = buf[:, :n]
        wt = None if w is None else torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32).ravel()).to(device)
        return cls(Xt, yt, wt)


_LOCAL_ONLY = 0


class single_rank:
    """``with single_rank(): gam.fit(...)`` -- fit (or predict) on this rank's rows only although a process group is
    initialised: no collective is issued inside the block (a side model fitted by one rank of a sharded job)"""

    def __enter__(self):
        global _LOCAL_ONLY
        _LOCAL_ONLY += 1
        return self

    def __exit__(self, *exc):
        global _LOCAL_ONLY
        _LOCAL_ONLY -= 1
        return False


def all_reduce_(t, op="sum"):
    """in-place all-reduce over the default process group when one is initialised"""
    import torch.distributed as dist

    if _LOCAL_ONLY == 0 and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        ops = {"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}
        dist.all_reduce(t, op=ops[op])
    return t


def world():
    """(rank, world size) of the default process group; (0, 1) without one"""
    import torch.distributed as dist

    if _LOCAL_ONLY == 0 and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_gather_rows(local, total_rows, lo, device):
    """rows [lo, lo + len(local)) of a (total_rows, width) host array live on this rank: returns the full array on
    every rank (grid-search candidates solved rank by rank).  One all-reduce(sum) of the zero-padded
    array: the message is small and every rank needs all of it."""
    full = torch.zeros((total_rows, local.shape[1]), dtype=torch.float64, device=device)
    if len(local):
        full[lo:lo + len(local)] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)).to(device)
    all_reduce_(full, "sum")
    return full.cpu().numpy()


def column_stats(data):
    """per-feature (min, max) and non-finite count of the (global) data: gen_edge_knots
    (utils.py:592-621) and the finite check of check_X (utils.py:180-182), on the device"""
    if "col_stats" not in data._memo:
        data._memo["col_stats"] = get_backend().column_stats(data)
    return data._memo["col_stats"]


def count_levels(data, feature, lo, hi):
    key = ("levels", feature)
    if key not in data._memo:
        data._memo[key] = get_backend().count_levels(data, feature, lo, hi)
    return data._memo[key]


def _cuda_column_stats(data):
    if not data.Xt.is_cuda:
        raise RuntimeError("pygam_b200 needs the data on a CUDA device; there is no CPU fallback")
    lib = _lib.load()
    F = data.n_features
    minmax = torch.empty(2 * F, dtype=torch.float64, device=data.device)
    bad = torch.empty(F, dtype=torch.int64, device=data.device)
    ws = torch.empty(int(lib.pgb_col_stats_ws_bytes(F)) // 8 + 1, dtype=torch.float64, device=data.device)
    _lib.check(lib.pgb_col_stats(_ptr(data.Xt), data.n, data.ld, F, _ptr(minmax), _ptr(bad), _ptr(ws),
                                 ws.numel() * 8, _stream()), "pgb_col_stats")
    mm = minmax.view(F, 2)
    lo = all_reduce_(mm[:, 0].contiguous(), "min")
    hi = all_reduce_(mm[:, 1].contiguous(), "max")
    bad = all_reduce_(bad, "sum")
    return np.stack([lo.cpu().numpy(), hi.cpu().numpy()], axis=1), bad.cpu().numpy()


def vector_range(y):
    """(min over the finite entries, max, number of non-finite entries) of a device vector in one pgb_col_stats pass
    (utils.check_y, utils.py:204-248, needs nothing else); local rows only"""
    if not y.is_cuda:  # the CPU test engine (tests/oracle_backend.py) only: the product has no CPU path
        if isinstance(get_backend(), CudaBackend):
            raise RuntimeError("pygam_b200 needs the data on a CUDA device; there is no CPU fallback")
        fin = torch.isfinite(y)
        yf = y[fin]
        return (float(yf.min()) if yf.numel() else np.inf, float(yf.max()) if yf.numel() else -np.inf, int((~fin).sum()))
    lib = _lib.load()
    out = torch.empty(3 + int(lib.pgb_col_stats_ws_bytes(1)) // 8 + 1, dtype=torch.float64, device=y.device)
    with torch.cuda.device(y.device):
        _lib.check(lib.pgb_col_stats(_ptr(y), y.numel(), y.numel(), 1, _ptr(out), _ptr(out[2:]), _ptr(out[3:]),
                                     (out.numel() - 3) * 8, _stream()), "pgb_col_stats")
    h = out[:3].cpu()
    return float(h[0]), float(h[1]), int(h[2:3].view(torch.int64)[0])


def _cuda_count_levels(data, feature, lo, hi):
    """number of distinct values of an integer-coded column (FactorTerm.compile, terms.py:1071)"""
    lib = _lib.load()
    rng = int(round(hi - lo)) + 1
    if not (1 <= rng <= 1 << 20):
        raise ValueError(f"categorical feature {feature} spans {rng} levels")
    present = torch.empty(rng + 1, dtype=torch.int32, device=data.device)
    _lib.check(lib.pgb_col_levels(_ptr(data.Xt[feature]), data.n, float(lo), rng, _ptr(present), _stream()),
               "pgb_col_levels")
    present = all_reduce_(present, "max")
    p = present.cpu().numpy()
    if p[rng] != 0:
        # non-integer levels: exact distinct count of the local shard; the ranks of a sharded fit would have to
        # exchange their level sets first
        if world()[1] > 1:
            raise NotImplementedError(f"categorical feature {feature} has non-integer levels: not supported with "
                                      "row shards on several ranks (code the levels as integers)")
        vals = torch.unique(data.Xt[feature])
        return int(vals.numel())
    return int(p[:rng].sum())


def descriptor_key(terms, n_features):
    """identity of the device model: everything the kernels see (not lam / penalties)"""
    return bytes(build_descriptor(terms, n_features))


def build_descriptor(terms, n_features):
    """flatten a compiled TermList into the pgb_term array of include/pgb.h"""
    arr = (_lib.PgbTerm * len(terms))()
    off = 0
    for i, t in enumerate(terms):
        d = arr[i]
        d.by = -1 if getattr(t, "by", None) is None else int(t.by)
        d.drop_first = 0
        d.coef_offset = off
        d.n_coefs = t.n_coefs
        if isinstance(t, Intercept):
            d.kind, d.n_marg = _lib.TERM_INTERCEPT, 0
        elif isinstance(t, LinearTerm):
            d.kind, d.n_marg = _lib.TERM_LINEAR, 1
            d.feature[0] = t.feature
            d.n_splines[0] = 1
        elif isinstance(t, TensorTerm):
            d.kind, d.n_marg = _lib.TERM_TENSOR, len(t._terms)
            for q, mg in enumerate(t._terms):
                _fill_marginal(d, q, mg)
        elif isinstance(t, FactorTerm):
            d.kind, d.n_marg = _lib.TERM_FACTOR, 1
            _fill_marginal(d, 0, t)
            d.drop_first = 1 if t.coding == "dummy" else 0
        elif isinstance(t, SplineTerm):
            d.kind, d.n_marg = _lib.TERM_SPLINE, 1
            _fill_marginal(d, 0, t)
        else:
            raise NotImplementedError(f"custom term type {type(t).__name__} has no CUDA basis evaluation")
        off += t.n_coefs
    return arr


def _fill_marginal(d, q, t):
    if isinstance(t, LinearTerm):  # l() marginal of te(): one column, the feature value (PGB_MARG_LINEAR)
        d.feature[q] = t.feature
        d.n_splines[q] = 1
        d.order[q] = -1
        d.periodic[q] = 0
        d.edge_lo[q], d.edge_hi[q] = 0.0, 1.0
        return
    if t.edge_knots_ is None:
        raise ValueError("term has not been compiled")
    d.feature[q] = t.feature
    d.n_splines[q] = t.n_splines
    d.order[q] = t.spline_order
    d.periodic[q] = 1 if t.basis == "cp" else 0
    d.edge_lo[q] = float(t.edge_knots_[0])
    d.edge_hi[q] = float(t.edge_knots_[-1])


class Engine:
    """compiled model + workspaces on one device"""

    def __init__(self, terms, dist_name, link_name, levels=1.0, expectile=None, n_features=None, device=None,
                 debug_opts=None):
        if not torch.cuda.is_available():
            raise RuntimeError("pygam_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dist_name not in _lib.DISTS or link_name not in _lib.LINKS:
            raise NotImplementedError("only the built-in distributions and links have CUDA implementations")
        self.m = terms.n_coefs
        self.n_features = int(n_features)
        desc = build_descriptor(terms, self.n_features)
        handle = C.c_void_p()
        args = (desc, len(terms), self.n_features, _lib.DISTS[dist_name], _lib.LINKS[link_name], float(levels),
                float(expectile) if expectile is not None else -1.0, self.device.index or 0)
        if debug_opts is None:
            _lib.check(self.lib.pgb_model_create(*args, C.byref(handle)), "pgb_model_create")
        else:  # tests only: the same model with the plan of its Gram pass constrained
            _lib.check(self.lib.pgb_debug_model_create(*args, C.byref(debug_opts), C.byref(handle)),
                       "pgb_debug_model_create")
        self.use_plan = True  # tests switch the stored row order off to exercise the in-kernel sort
        self.handle = handle
        info = _lib.PgbModelInfo()
        _lib.check(self.lib.pgb_model_get_info(self.handle, C.byref(info)), "pgb_model_get_info")
        self.info = info
        m = self.m
        f64 = dict(dtype=torch.float64, device=self.device)
        self.out = torch.zeros(m * m + m + _lib.GS_COUNT, **f64)
        self.beta = torch.zeros(m, **f64)
        self.beta_prev = torch.zeros(m, **f64)
        self.EtE = torch.zeros(m * m, **f64)
        self.EtE_null = torch.zeros(m * m, **f64)
        self.A0 = torch.zeros(m * m, **f64)  # penalty in the null-space basis (pgb_solve_prepare)
        self._prep_ws = torch.zeros(m * m, **f64)
        self._has_null_part = False
        self.sol = torch.zeros(m + _lib.SOLVE_OUT_EXTRA, **f64)
        self.cov = None
        self.stat_scalars = torch.zeros(_lib.SS_COUNT, **f64)
        self._gram_ws = None
        self._plan_key = None
        self._plan_data = None
        self._plan_passes = 0
        self._ws_has_order = True
        self._stats_ws = torch.empty(int(self.lib.pgb_stats_ws_bytes(self.handle)) // 8 + 1, **f64)
        self._solve_ws = torch.empty(int(self.lib.pgb_solve_ws_bytes(m, 1)) // 8 + 1, **f64)
        self._chol_ws = None
        self._chol_info = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.hv = None
        self.tau = None
        self.p = 0
        self.set_null_space(terms.null_vectors())

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.pgb_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- null-space reflectors (host, O(m p^2)) ----
    def set_null_space(self, N):
        """Householder QR of the analytic null vectors N (m, p): Q = H_0 ... H_{p-1}, first p
        columns of Q span N.  Vectors and taus go to the device for pgb_solve."""
        m, p = N.shape
        self.p = p
        if p == 0:
            self.hv = self.tau = None
            return
        A = np.array(N, dtype=np.float64)
        V = np.zeros((p, m))
        tau = np.zeros(p)
        for q in range(p):
            x = A[q:, q].copy()
            nx = np.linalg.norm(x)
            if nx == 0.0:
                continue
            alpha = -np.copysign(nx, x[0])
            v = x
            v[0] -= alpha
            vn2 = v @ v
            if vn2 == 0.0:
                continue
            V[q, q:] = v
            tau[q] = 2.0 / vn2
            A[q:, q:] -= tau[q] * np.outer(v, v @ A[q:, q:])
        self.hv = torch.from_numpy(V.ravel()).to(self.device)
        self.tau = torch.from_numpy(tau).to(self.device)

    # ---- passes ----
    def _ensure_gram_ws(self, n):
        need = int(self.lib.pgb_gram_ws_bytes(self.handle, n)) // 8 + 1
        if self._gram_ws is None or self._gram_ws.numel() < need:
            self._plan_key = None
            self._gram_ws = None
            try:
                self._gram_ws = torch.empty(need, dtype=torch.float64, device=self.device)
                self._ws_has_order = True
            except torch.cuda.OutOfMemoryError:
                # no room for the stored row order (2 bytes per row and term pair): sort inside the kernel
                small = int(self.lib.pgb_gram_ws_bytes_streaming(self.handle, n)) // 8 + 1
                self._gram_ws = torch.empty(small, dtype=torch.float64, device=self.device)
                self._ws_has_order = False
        return self._gram_ws

    def _planned(self, data, ws):
        """What does not depend on the coefficients (knot cells of the rows and their order, all-spline models)
        is prepared once per resident shard: the reference builds its model matrix once per fit, too
        (pygam.py:741).  Returns the mode flag for pgb_gram_pass."""
        if self.info.cell_roles == 0 or not self._ws_has_order or not self.use_plan:
            return 0
        key = (data.Xt.data_ptr(), data.n, ws.data_ptr(), data.Xt._version)
        if self._plan_key != key:
            _lib.check(self.lib.pgb_gram_prepare(self.handle, _ptr(data.Xt), data.n, data.ld, _ptr(ws), ws.numel() * 8,
                                                 _stream()), "pgb_gram_prepare")
            self._plan_key = key
            self._plan_data = data.Xt  # keeps the tensor (and its address) alive as long as the plan
            self._plan_passes = 0
        # a shard that keeps being passed over (a long fit, a grid search on the same rows) gets the second step of
        # the preparation once: it costs about three passes and takes 13 % off every later one
        self._plan_passes += 1
        if self._plan_passes == self.REFINE_AFTER_PASSES:
            self.refine_order(data)
        return _lib.MODE_PLANNED

    # pgb_gram_refine costs 1.5 (config 3) to 3.5 (config 2) Gram passes and saves 1-11 % of every later one
    # (profiles/r2_kbench.log): it pays after 28-140 passes.  A shard that has been passed over 50 times (a grid
    # search or bootstrap refits, not a single fit) is expected to be passed over as often again.
    REFINE_AFTER_PASSES = 50

    def refine_order(self, data):
        """pgb_gram_refine on the prepared order of `data` (once per preparation)"""
        ws = self._gram_ws
        if ws is None or self._plan_key is None or self._plan_passes < 0:
            return
        _lib.check(self.lib.pgb_gram_refine(self.handle, data.n, _ptr(ws), ws.numel() * 8, _stream()), "pgb_gram_refine")
        self._plan_passes = -(1 << 40)  # refined: never again for this preparation

    def gram_pass(self, data, coef=None, mode=_lib.MODE_IRLS, reduce=True):
        """G = X^T W^2 X, r = X^T W^2 z and the iteration scalars at `coef` (device tensors).
        Returns the packed buffer [G | r | scalars] (all-reduced over row shards)."""
        ws = self._ensure_gram_ws(data.n)
        if coef is not None:
            self.beta.copy_(torch.as_tensor(coef, dtype=torch.float64))
        mode = mode | self._planned(data, ws)
        _lib.check(self.lib.pgb_gram_pass(self.handle, mode, _ptr(data.Xt), data.n, data.ld, _ptr(data.y),
                                          _ptr(data.w), _ptr(self.beta), _ptr(self.out), _ptr(ws),
                                          ws.numel() * 8, _stream()), "pgb_gram_pass")
        if reduce:
            all_reduce_(self.out, "sum")
        return self.out

    def gram_scalars(self):
        return self.out[self.m * self.m + self.m:].cpu().numpy()

    def set_penalty(self, EtE, EtE_null=None):
        """penalty of the m x m system as EtE + EtE_null; EtE_null holds the components that
        annihilate the analytic null vectors structurally (see pgb_solve in include/pgb.h)"""
        self.EtE.copy_(torch.from_numpy(np.ascontiguousarray(EtE, dtype=np.float64).ravel()))
        self._has_null_part = EtE_null is not None
        if EtE_null is not None:
            self.EtE_null.copy_(torch.from_numpy(np.ascontiguousarray(EtE_null, dtype=np.float64).ravel()))
        # the change of basis of the penalty is done once here, not in every PIRLS iteration
        m = self.m
        _lib.check(self.lib.pgb_solve_prepare(m, 1, _ptr(self.EtE),
                                              _ptr(self.EtE_null) if self._has_null_part else C.c_void_p(0),
                                              _ptr(self.hv), _ptr(self.tau), self.p, _ptr(self.A0), _ptr(self._prep_ws),
                                              self._prep_ws.numel() * 8, _stream()), "pgb_solve_prepare")

    def penalty_host(self):
        """the penalty of the last set_penalty, E^T E = S + P (+ C), as an (m, m) host array"""
        E = self.EtE.view(self.m, self.m).cpu().numpy().copy()
        if self._has_null_part:
            E += self.EtE_null.view(self.m, self.m).cpu().numpy()
        return E

    def chol_ok(self, A):
        """True when A (host, m x m) has a Cholesky factor: utils.cholesky (utils.py:77-91)"""
        m = self.m
        need = int(self.lib.pgb_chol_ws_bytes(m)) // 8 + 1
        if self._chol_ws is None:
            self._chol_ws = torch.empty(need + m * m, dtype=torch.float64, device=self.device)
        Ad = self._chol_ws[need:need + m * m]
        Ad.copy_(torch.from_numpy(np.ascontiguousarray(A, dtype=np.float64).ravel()))
        _lib.check(self.lib.pgb_chol_check(m, _ptr(Ad), _ptr(self._chol_info), _ptr(self._chol_ws), need * 8,
                                           _stream()), "pgb_chol_check")
        return int(self._chol_info.item()) == 0

    def chol_ok_batch(self, A):
        """utils.cholesky's test (utils.py:77-91) for K matrices at once: (K, m, m) host array -> bool (K,)"""
        m = self.m
        A = np.ascontiguousarray(A, dtype=np.float64).reshape(-1, m, m)
        K = A.shape[0]
        ok = np.zeros(K, dtype=bool)
        chunk = max(1, min(K, (1 << 28) // (m * m * 8)))
        for k0 in range(0, K, chunk):
            nb = min(chunk, K - k0)
            Ad = torch.from_numpy(A[k0:k0 + nb].reshape(-1)).to(self.device)
            info = torch.zeros(nb, dtype=torch.int32, device=self.device)
            ws = torch.empty(nb * m * m, dtype=torch.float64, device=self.device)
            _lib.check(self.lib.pgb_chol_check_batch(m, nb, _ptr(Ad), _ptr(info), _ptr(ws), ws.numel() * 8, _stream()),
                       "pgb_chol_check_batch")
            ok[k0:k0 + nb] = info.cpu().numpy() == 0
        return ok

    def solve(self, coef_prev=None, want_edof=False, want_cov=False, out=None):
        """beta = (G + EtE)^-1 r with the Gram data currently in self.out; returns host dict"""
        m = self.m
        flags = (_lib.SOLVE_EDOF if want_edof else 0) | (_lib.SOLVE_COV if want_cov else 0) | _lib.SOLVE_PREPARED
        if want_cov and self.cov is None:
            self.cov = torch.empty(m * m, dtype=torch.float64, device=self.device)
        prev = None
        if coef_prev is not None:
            self.beta_prev.copy_(torch.as_tensor(coef_prev, dtype=torch.float64))
            prev = self.beta_prev
        G = self.out
        r = self.out[m * m:]
        _lib.check(self.lib.pgb_solve(m, 1, _ptr(G), _ptr(r), 1, _ptr(self.A0), C.c_void_p(0),
                                      _ptr(self.hv), _ptr(self.tau),
                                      self.p, _ptr(prev), 1, flags, _ptr(self.sol),
                                      _ptr(self.cov) if want_cov else C.c_void_p(0), _ptr(self._solve_ws),
                                      self._solve_ws.numel() * 8, _stream()), "pgb_solve")
        s = self.sol.cpu().numpy()
        res = dict(coef=s[:m].copy(), edof=float(s[m]), diff=float(s[m + 1]), info=int(s[m + 2]) if np.isfinite(s[m + 2]) else 2,
                   logdet=float(s[m + 3]))
        if want_cov:
            res["cov"] = self.cov.view(m, m).cpu().numpy().copy()
        return res

    def solve_truncated(self, rank, coef_prev=None, want_cov=False):
        """fewer kept rows than coefficients: the reference's SVD cropped to `rank` singular triplets
        (pygam.py:789-799) = the `rank` largest eigenpairs of G + EtE, with the Gram data in self.out and the
        penalty of the last set_penalty; same result dict as solve()"""
        m = self.m
        if want_cov and self.cov is None:
            self.cov = torch.empty(m * m, dtype=torch.float64, device=self.device)
        prev = None
        if coef_prev is not None:
            self.beta_prev.copy_(torch.as_tensor(coef_prev, dtype=torch.float64))
            prev = self.beta_prev
        need = int(self.lib.pgb_solve_truncated_ws_bytes(m)) // 8 + 1
        if getattr(self, "_trunc_ws", None) is None:
            self._trunc_ws = torch.empty(need, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.pgb_solve_truncated(m, int(rank), _ptr(self.out), _ptr(self.out[m * m:]), _ptr(self.EtE),
                                                _ptr(self.EtE_null) if self._has_null_part else C.c_void_p(0), _ptr(prev),
                                                _lib.SOLVE_COV if want_cov else 0, _ptr(self.sol),
                                                _ptr(self.cov) if want_cov else C.c_void_p(0), _ptr(self._trunc_ws),
                                                self._trunc_ws.numel() * 8, _stream()), "pgb_solve_truncated")
        s = self.sol.cpu().numpy()
        res = dict(coef=s[:m].copy(), edof=float(s[m]), diff=float(s[m + 1]),
                   info=int(s[m + 2]) if np.isfinite(s[m + 2]) else 2, logdet=0.0)
        if not np.isfinite(res["coef"]).all():
            res["info"] = 2
        if want_cov:
            res["cov"] = self.cov.view(m, m).cpu().numpy().copy()
        return res

    def solve_batch(self, EtE, EtE_null=None, want_cov=False, max_bytes=1 << 28):
        """K penalised systems that share the Gram data currently in self.out (grid-search candidates of a model
        whose weights do not depend on the coefficients): one CTA per system, pgb_solve(nb = K).
        EtE, EtE_null: (K, m, m) host arrays.  Returns coef (K, m), edof (K,), info (K,) [, cov (K, m, m)]."""
        m = self.m
        EtE = np.ascontiguousarray(EtE, dtype=np.float64).reshape(-1, m, m)
        K = EtE.shape[0]
        if EtE_null is not None:
            EtE_null = np.ascontiguousarray(EtE_null, dtype=np.float64).reshape(K, m, m)
        per = m * m * 8 * (4 + (1 if want_cov else 0))
        chunk = max(1, min(K, int(max_bytes // per)))
        f64 = dict(dtype=torch.float64, device=self.device)
        coef = np.empty((K, m))
        edof = np.empty(K)
        info = np.empty(K, dtype=np.int64)
        cov = np.empty((K, m, m)) if want_cov else None
        flags = _lib.SOLVE_EDOF | _lib.SOLVE_PREPARED | (_lib.SOLVE_COV if want_cov else 0)
        ws = torch.empty(int(self.lib.pgb_solve_ws_bytes(m, chunk)) // 8 + 1, **f64)
        prep_ws = torch.empty(chunk * m * m, **f64)
        A0 = torch.empty(chunk * m * m, **f64)
        out = torch.empty(chunk * (m + _lib.SOLVE_OUT_EXTRA), **f64)
        covd = torch.empty(chunk * m * m, **f64) if want_cov else None
        G = self.out
        r = self.out[m * m:]
        for k0 in range(0, K, chunk):
            nb = min(chunk, K - k0)
            E = torch.from_numpy(EtE[k0:k0 + nb].reshape(-1)).to(self.device)
            En = None if EtE_null is None else torch.from_numpy(EtE_null[k0:k0 + nb].reshape(-1)).to(self.device)
            _lib.check(self.lib.pgb_solve_prepare(m, nb, _ptr(E), _ptr(En), _ptr(self.hv), _ptr(self.tau), self.p, _ptr(A0),
                                                  _ptr(prep_ws), prep_ws.numel() * 8, _stream()), "pgb_solve_prepare")
            _lib.check(self.lib.pgb_solve(m, nb, _ptr(G), _ptr(r), 1, _ptr(A0), C.c_void_p(0), _ptr(self.hv),
                                          _ptr(self.tau), self.p, C.c_void_p(0), 1, flags, _ptr(out), _ptr(covd),
                                          _ptr(ws), ws.numel() * 8, _stream()), "pgb_solve")
            o = out[:nb * (m + _lib.SOLVE_OUT_EXTRA)].view(nb, m + _lib.SOLVE_OUT_EXTRA).cpu().numpy()
            coef[k0:k0 + nb] = o[:, :m]
            edof[k0:k0 + nb] = o[:, m]
            inf = o[:, m + 2]
            info[k0:k0 + nb] = np.where(np.isfinite(inf), inf, 2).astype(np.int64)
            if want_cov:
                cov[k0:k0 + nb] = covd[:nb * m * m].view(nb, m, m).cpu().numpy()
        res = dict(coef=coef, edof=edof, info=info)
        if want_cov:
            res["cov"] = cov
        return res

    def stats_pass(self, data, coef, scale=-1.0, ybar=0.0, poisson_round=False, want_mu=False, want_lp=False,
                   reduce=True):
        """lp / mu and the statistics sums on ALL rows (pygam.py:1031-1056, 423-450)"""
        self.beta.copy_(torch.as_tensor(coef, dtype=torch.float64))
        mu = torch.empty(data.n, dtype=torch.float64, device=self.device) if want_mu else None
        lp = torch.empty(data.n, dtype=torch.float64, device=self.device) if want_lp else None
        _lib.check(self.lib.pgb_stats_pass(self.handle, _ptr(data.Xt), data.n, data.ld, _ptr(data.y), _ptr(data.w),
                                           _ptr(self.beta), float(scale), float(ybar), 1 if poisson_round else 0,
                                           _ptr(self.stat_scalars), _ptr(lp), _ptr(mu), _ptr(self._stats_ws),
                                           self._stats_ws.numel() * 8, _stream()), "pgb_stats_pass")
        if reduce:
            all_reduce_(self.stat_scalars, "sum")
        return self.stat_scalars.cpu().numpy().copy(), lp, mu

    def interval_pass(self, data, coef, cov=None, term=-1):
        """lp of one term (or all, term = -1) and var_i = x_i^T cov x_i per row (pygam.py:1399-1404);
        returns device tensors (lp, var); var is None without cov"""
        self.beta.copy_(torch.as_tensor(coef, dtype=torch.float64))
        lp = torch.empty(data.n, dtype=torch.float64, device=self.device)
        var = covd = None
        if cov is not None:
            covd = torch.from_numpy(np.ascontiguousarray(cov, dtype=np.float64)).to(self.device)
            var = torch.empty(data.n, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.pgb_interval_pass(self.handle, _ptr(data.Xt), data.n, data.ld, _ptr(self.beta), _ptr(covd),
                                              int(term), _ptr(lp), _ptr(var), _stream()), "pgb_interval_pass")
        return lp, var

    def model_matrix(self, data):
        out = torch.empty(data.n * self.m, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.pgb_model_matrix(self.handle, _ptr(data.Xt), data.n, data.ld, _ptr(out), _stream()),
                   "pgb_model_matrix")
        return out.view(data.n, self.m)


class CudaBackend:
    """the product backend: hand-written sm_100a kernels behind the C ABI"""

    name = "cuda"

    def default_device(self):
        if not torch.cuda.is_available():
            raise RuntimeError("pygam_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())

    def column_stats(self, data):
        return _cuda_column_stats(data)

    def count_levels(self, data, feature, lo, hi):
        return _cuda_count_levels(data, feature, lo, hi)

    def make_engine(self, terms, dist_name, link_name, levels, expectile, n_features, device):
        return Engine(terms, dist_name, link_name, levels=levels, expectile=expectile, n_features=n_features,
                      device=device)


_BACKEND = CudaBackend()


def get_backend():
    return _BACKEND


def swap_backend_for_tests(backend):
    """TEST HOOK, never called by the product: the CPU test-suite swaps in an oracle-backed engine
    (tests/oracle_backend.py) to exercise the host logic -- PIRLS driver, grid search, sharding -- without a GPU.
    The product has exactly one backend, the CUDA engine above."""
    global _BACKEND
    old = _BACKEND
    _BACKEND = backend
    return old
