"""TEST INFRASTRUCTURE: a CPU stand-in for pygam_b200.engine.Engine built on the oracle.

It lets the CPU test-suite exercise the host logic of the product (PIRLS driver, penalty /
constraint handling, grid search, statistics assembly, row sharding with an all-reduce)
without a GPU.  The per-row arithmetic comes from oracle/gam_oracle.py; the m x m solve is a
numpy mirror of the algorithm of pygam_b200/csrc/pgb_solve.cu (Householder deflation of the
analytic null space, Cholesky, edof / cov from the factor), so the tests also pin that
algorithm's parity with the reference's QR+SVD method.  Never imported by the product.
"""
import numpy as np
import scipy.linalg
import torch

from oracle import gam_oracle as orc
from pygam_b200 import _lib
from pygam_b200.engine import all_reduce_
from pygam_b200.terms import FactorTerm, Intercept, LinearTerm, SplineTerm, TensorTerm


def _oracle_marg(t):
    if isinstance(t, LinearTerm):
        return dict(kind="l", feature=t.feature)
    return dict(kind="s", feature=t.feature, n_splines=t.n_splines, spline_order=t.spline_order,
                basis=t.basis, by=t.by, dtype=t.dtype, edge_knots=np.asarray(t.edge_knots_, float),
                penalties=["none"], lam=[0.0], constraints=[None])


def oracle_terms(terms):
    out = []
    for t in terms:
        if isinstance(t, Intercept):
            out.append(dict(kind="intercept"))
        elif isinstance(t, LinearTerm):
            out.append(dict(kind="l", feature=t.feature))
        elif isinstance(t, TensorTerm):
            out.append(dict(kind="te", marginals=[_oracle_marg(q) for q in t._terms], by=t.by))
        elif isinstance(t, FactorTerm):
            d = _oracle_marg(t)
            d.update(kind="f", coding=t.coding)
            out.append(d)
        elif isinstance(t, SplineTerm):
            out.append(_oracle_marg(t))
        else:
            raise NotImplementedError
    return out


def householder(N):
    m, p = N.shape
    A = np.array(N, dtype=np.float64)
    V = np.zeros((p, m))
    tau = np.zeros(p)
    for q in range(p):
        x = A[q:, q].copy()
        nx = np.linalg.norm(x)
        if nx == 0.0:
            continue
        x[0] -= -np.copysign(nx, x[0])
        V[q, q:] = x
        tau[q] = 2.0 / (x @ x)
        A[q:, q:] -= tau[q] * np.outer(x, x @ A[q:, q:])
    return V, tau


def solve_mirror(G, r, EtE, V, tau, want_cov=False, EtE_null=None):
    """numpy mirror of solve_kernel (pgb_solve.cu)"""
    m = len(r)
    p = len(tau)
    Q = np.eye(m)
    for q in range(p):
        Q = Q - tau[q] * np.outer(Q @ V[q], V[q])  # Q = H_0 H_1 ... H_{p-1}
    Gt = Q.T @ G @ Q
    Gt[:p, :] = 0.0
    Gt[:, :p] = 0.0
    A = Q.T @ EtE @ Q + Gt
    if EtE_null is not None:
        En = Q.T @ EtE_null @ Q
        En[:p, :] = 0.0
        En[:, :p] = 0.0
        A = A + En
    x = Q.T @ r
    x[:p] = 0.0
    if not (np.isfinite(G).all() and np.isfinite(EtE).all() and np.isfinite(r).all()):
        return dict(info=2)
    try:
        L = scipy.linalg.cholesky(A, lower=True)
    except scipy.linalg.LinAlgError:
        return dict(info=1)
    beta_t = scipy.linalg.cho_solve((L, True), x)
    out = dict(info=0, coef=Q @ beta_t)
    Z = scipy.linalg.solve_triangular(L, Gt, lower=True)
    Z = scipy.linalg.solve_triangular(L, Z.T, lower=True)
    out["edof"] = float(np.trace(Z))
    if want_cov:
        C = scipy.linalg.solve_triangular(L, Z, lower=True, trans="T")
        C = scipy.linalg.solve_triangular(L, C.T, lower=True, trans="T")
        out["cov"] = Q @ C @ Q.T
    return out


class OracleEngine:
    def __init__(self, terms, dist_name, link_name, levels, expectile, n_features, device):
        self.m = terms.n_coefs
        self.oterms = oracle_terms(terms)
        self.term_coefs = [t.n_coefs for t in terms]
        self.fam = dict(dist=dist_name, link=link_name, levels=levels, expectile=expectile, scale=None)
        self.V, self.tau = householder(terms.null_vectors())
        m = self.m
        self.out = torch.zeros(m * m + m + _lib.GS_COUNT, dtype=torch.float64)
        self.EtE = np.zeros((m, m))
        self.EtE_null = None
        self._mm_cache = None

    def _modelmat(self, data):
        if self._mm_cache is None or self._mm_cache[0] is not data.Xt:
            X = data.Xt.numpy().T
            self._mm_cache = (data.Xt, orc.model_matrix(self.oterms, X).tocsr())
        return self._mm_cache[1]

    def gram_pass(self, data, coef=None, mode=_lib.MODE_IRLS, reduce=True):
        B = self._modelmat(data)
        m = self.m
        y = data.y.numpy()
        w = np.ones_like(y) if data.w is None else data.w.numpy()  # float32, as in the reference
        sc = np.zeros(_lib.GS_COUNT)
        if mode == _lib.MODE_INIT:
            yy = y.copy()
            yy[yy == 0] += 0.01
            yy[yy == 1] -= 0.01
            z = orc.link_fn(self.fam["link"], yy, self.fam["levels"])
            om = np.ones_like(y)
            sc[_lib.GS_NUSED] = len(y)
            sc[_lib.GS_NONFINITE] = np.sum(~np.isfinite(z))
        else:
            coef = np.asarray(coef, float)
            lp = B.dot(coef)
            with np.errstate(all="ignore"):
                mu = orc.link_inv(self.fam["link"], lp, self.fam["levels"])
                wd = orc.irls_weight(self.fam, mu, w, y)
                mask = (np.abs(wd) >= np.sqrt(orc.EPS)) * np.isfinite(wd)
                z = lp + (y - mu) * orc.link_grad(self.fam["link"], mu, self.fam["levels"])
            om = np.where(mask, wd ** 2, 0.0)
            z = np.where(mask, z, 0.0)
            sc[_lib.GS_NUSED] = mask.sum()
            if mask.any():
                with np.errstate(all="ignore"):
                    sc[_lib.GS_DEV] = orc.deviance_rows(self.fam["dist"], y[mask], mu[mask], self.fam["levels"],
                                                        None, scaled=False).sum()
                sc[_lib.GS_NCORRECT] = np.sum(y[mask] == (mu[mask] > 0.5))
        sc[_lib.GS_NROWS] = len(y)
        WB = B.multiply(om[:, None]).tocsr()
        G = (B.T @ WB).toarray()
        r = B.T @ (om * z)
        self.out[: m * m] = torch.from_numpy(G.ravel())
        self.out[m * m: m * m + m] = torch.from_numpy(np.asarray(r).ravel())
        self.out[m * m + m:] = torch.from_numpy(sc)
        if reduce:
            all_reduce_(self.out, "sum")
        return self.out

    def gram_scalars(self):
        return self.out[self.m * self.m + self.m:].numpy().copy()

    def set_penalty(self, EtE, EtE_null=None):
        self.EtE = np.array(EtE, dtype=np.float64)
        self.EtE_null = None if EtE_null is None else np.array(EtE_null, dtype=np.float64)

    def penalty_host(self):
        return self.EtE + (0.0 if self.EtE_null is None else self.EtE_null)

    def chol_ok(self, A):
        try:
            scipy.linalg.cholesky(A, lower=False)
            return True
        except scipy.linalg.LinAlgError:
            return False

    def chol_ok_batch(self, A):
        return np.array([self.chol_ok(a) for a in A], dtype=bool)

    def solve(self, coef_prev=None, want_edof=False, want_cov=False):
        m = self.m
        o = self.out.numpy()
        G = o[: m * m].reshape(m, m)
        r = o[m * m: m * m + m]
        res = solve_mirror(G, r, self.EtE, self.V, self.tau, want_cov=want_cov, EtE_null=self.EtE_null)
        if res["info"] != 0:
            return dict(info=res["info"], coef=np.full(m, np.nan), edof=np.nan, diff=np.nan)
        res["diff"] = np.nan
        if coef_prev is not None:
            res["diff"] = np.linalg.norm(np.asarray(coef_prev) - res["coef"]) / np.linalg.norm(res["coef"])
        return res

    def solve_truncated(self, rank, coef_prev=None, want_cov=False):
        """numpy mirror of pgb_solve_truncated: the `rank` largest eigenpairs of G + EtE"""
        m = self.m
        o = self.out.numpy()
        G = o[: m * m].reshape(m, m)
        r = o[m * m: m * m + m]
        A = G + self.EtE + (0.0 if self.EtE_null is None else self.EtE_null)
        lam, V = np.linalg.eigh(0.5 * (A + A.T))
        keep = np.argsort(-lam, kind="stable")[:rank]
        Vk, lk = V[:, keep], lam[keep]
        coef = Vk @ ((Vk.T @ r) / lk)
        W = (Vk.T @ G @ Vk) / np.outer(lk, lk)
        res = dict(info=0, coef=coef, edof=float(np.sum(np.diag(W) * lk)), diff=np.nan)
        if coef_prev is not None:
            res["diff"] = np.linalg.norm(np.asarray(coef_prev) - coef) / np.linalg.norm(coef)
        if want_cov:
            res["cov"] = Vk @ W @ Vk.T
        return res

    def solve_batch(self, EtE, EtE_null=None, want_cov=False, max_bytes=None):
        m = self.m
        o = self.out.numpy()
        G = o[: m * m].reshape(m, m)
        r = o[m * m: m * m + m]
        K = len(EtE)
        coef, edof, info = np.full((K, m), np.nan), np.full(K, np.nan), np.zeros(K, dtype=np.int64)
        cov = np.full((K, m, m), np.nan) if want_cov else None
        for k in range(K):
            res = solve_mirror(G, r, EtE[k], self.V, self.tau, want_cov=want_cov,
                               EtE_null=None if EtE_null is None else EtE_null[k])
            info[k] = res["info"]
            if res["info"] == 0:
                coef[k], edof[k] = res["coef"], res["edof"]
                if want_cov:
                    cov[k] = res["cov"]
        out = dict(coef=coef, edof=edof, info=info)
        if want_cov:
            out["cov"] = cov
        return out

    def stats_pass(self, data, coef, scale=-1.0, ybar=0.0, poisson_round=False, want_mu=False, want_lp=False,
                   reduce=True):
        B = self._modelmat(data)
        lp = B.dot(np.asarray(coef, float))
        f = self.fam
        with np.errstate(all="ignore"):
            mu = orc.link_inv(f["link"], lp, f["levels"])
        sc = np.zeros(_lib.SS_COUNT)
        if data.y is not None:
            y = data.y.numpy()
            w = np.ones_like(y) if data.w is None else data.w.numpy().astype(np.float64)
            with np.errstate(all="ignore"):
                d = orc.deviance_rows(f["dist"], y, mu, f["levels"], None, scaled=False)
                sc[_lib.SS_N] = len(y)
                sc[_lib.SS_WDEV] = (w * d).sum()
                sc[_lib.SS_PEARSON] = np.sum(w * (y - mu) ** 2 / orc.variance(f["dist"], mu, f["levels"]))
                sc[_lib.SS_DEV] = d.sum()
                sc[_lib.SS_NCORRECT] = np.sum(y == (mu > 0.5))
                sc[_lib.SS_SUMY] = y.sum()
                sc[_lib.SS_SUMW] = w.sum()
                if scale > 0:
                    yl = np.round(y * w).astype("int") if poisson_round else y
                    nm = np.full_like(y, ybar)
                    sc[_lib.SS_LOGLIK] = orc.log_pdf_rows(f["dist"], yl, mu, w, f["levels"], scale).sum()
                    sc[_lib.SS_NULL_WDEV] = (w * orc.deviance_rows(f["dist"], y, nm, f["levels"], None, False)).sum()
                    sc[_lib.SS_NULL_LOGLIK] = orc.log_pdf_rows(f["dist"], yl, nm, w, f["levels"], scale).sum()
        t = torch.from_numpy(sc)
        if reduce:
            all_reduce_(t, "sum")
        return t.numpy().copy(), (torch.from_numpy(lp) if want_lp else None), (torch.from_numpy(mu) if want_mu else None)

    def model_matrix(self, data):
        return torch.from_numpy(self._modelmat(data).toarray())

    def interval_pass(self, data, coef, cov=None, term=-1):
        """numpy restatement of GAM._get_quantiles' n-sized part (pygam.py:1399-1404)"""
        B = self._modelmat(data).toarray()
        if term >= 0:
            start = sum(self.term_coefs[:term])
            idx = np.arange(start, start + self.term_coefs[term])
        else:
            idx = np.arange(self.m)
        Bt = B[:, idx]
        lp = Bt.dot(np.asarray(coef)[idx])
        var = None
        if cov is not None:
            c = np.asarray(cov)[idx][:, idx]
            var = torch.from_numpy((Bt.dot(c) * Bt).sum(axis=1))
        return torch.from_numpy(lp), var


class OracleBackend:
    name = "oracle-test"

    def default_device(self):
        return torch.device("cpu")

    def column_stats(self, data):
        X = data.Xt
        lo = all_reduce_(X.min(dim=1).values.clone(), "min")
        hi = all_reduce_(X.max(dim=1).values.clone(), "max")
        bad = all_reduce_((~torch.isfinite(X)).sum(dim=1), "sum")
        return np.stack([lo.numpy(), hi.numpy()], axis=1), bad.numpy()

    def count_levels(self, data, feature, lo, hi):
        rng = int(round(hi - lo)) + 1
        present = torch.zeros(rng, dtype=torch.int32)
        idx = (data.Xt[feature] - lo).round().long()
        present[idx] = 1
        all_reduce_(present, "max")
        return int(present.sum())

    def make_engine(self, terms, dist_name, link_name, levels, expectile, n_features, device):
        return OracleEngine(terms, dist_name, link_name, levels, expectile, n_features, device)
