"""GAM model API on the B200 PIRLS engine.

Public surface mirrors ``/root/reference/pygam/pygam.py``: ``GAM`` (91-2368) and the family
front-ends ``LinearGAM`` (2371-2506), ``LogisticGAM`` (2509-2682), ``PoissonGAM`` (2685-3046),
``GammaGAM`` (3049-3165), ``InvGaussGAM`` (3168-3284), ``ExpectileGAM`` (3287-3551) with
``fit / predict / predict_mu / gridsearch``, fitted attributes ``coef_``, ``statistics_``,
``logs_`` and the reference's exceptions.  The solver behind them (``_pirls``, pygam.py:715-822)
is re-designed: a fused on-the-fly-basis Gram pass on the GPU, an all-reduce over row shards and
a null-space-aware m x m solve; nothing n-sized is computed on the host.
"""

import copy
import warnings
from collections import OrderedDict

import numpy as np
import scipy.linalg
import scipy.stats
import torch

from . import _lib
from .callbacks import builtin_name
from .engine import (DeviceData, all_gather_rows, all_reduce_, column_stats, count_levels, descriptor_key,
                     get_backend, single_rank, vector_range, world)
from .terms import FactorTerm, Intercept, SplineTerm, Term, TermList

EPS = np.finfo(np.float64).eps  # pygam/utils.py:16


class NotPositiveDefiniteError(ValueError):
    """pygam/utils.py:22-23"""


class OptimizationError(ValueError):
    """pygam/utils.py:26-27"""


# Host-side Distribution / Link objects with the reference's methods (distributions.py, links.py) for the small
# arrays a caller handles himself -- gam.distribution.V(mu), gam.link.mu(lp, dist), sampling; the n-sized link,
# variance, deviance and log-likelihood work of a fit or a predict runs in the kernels (pgb_device.cuh), selected by
# `name`.  Only the built-in families have CUDA implementations: custom Python objects raise NotImplementedError.
# pygam_b200.distributions / pygam_b200.links hold the reference's class names (NormalDist, LogitLink, ...).
def _ylogydu(a, u):
    """utils.py:773-789: y log(y / u), 0 where y == 0"""
    with np.errstate(all="ignore"):
        return np.where(a != 0, a * np.log(np.where(a != 0, a, 1.0) / u), 0.0)


class Distribution:
    def __init__(self, name, scale=None, levels=1):
        self.name = name
        self._known_scale = name in ("binomial", "poisson") or scale is not None
        self.scale = 1.0 if name in ("binomial", "poisson") else scale
        self.levels = 1 if levels is None else levels

    def __repr__(self):  # as an unfitted reference GAM shows its distribution argument
        return repr(self.name)

    def V(self, mu, weights=None):
        """variance function over the weights (distributions.py:134-160, 250-265, 355-370, 452-467, 555-570, 23-30)"""
        mu = np.asarray(mu, dtype=np.float64)
        v = {"normal": lambda: np.ones_like(mu), "binomial": lambda: mu * (1 - mu / self.levels), "poisson": lambda: mu,
             "gamma": lambda: mu ** 2}.get(self.name, lambda: mu ** 3)()
        return v if weights is None else v / np.asarray(weights)

    def deviance(self, y, mu, scaled=True, weights=None):
        """deviance of every row (distributions.py:162-185, 267-291, 372-397, 469-494, 572-597, decorator 13-20)"""
        name, lv = self.name, float(self.levels)
        y, mu = np.asarray(y, dtype=np.float64), np.asarray(mu, dtype=np.float64)
        with np.errstate(all="ignore"):
            if name == "normal":
                dev = (y - mu) ** 2
            elif name == "binomial":
                dev = 2 * (_ylogydu(y, mu) + _ylogydu(lv - y, lv - mu))
            elif name == "poisson":
                dev = 2 * (_ylogydu(y, mu) - (y - mu))
            elif name == "gamma":
                dev = 2 * ((y - mu) / mu - np.log(y / mu))
            else:
                dev = ((y - mu) ** 2) / (mu ** 2 * y)
        if scaled and self.scale:
            dev = dev / (self.scale ** 2 if name == "normal" else self.scale)
        if weights is not None:
            dev = dev * np.asarray(weights)
        return dev

    def log_pdf(self, y, mu, weights=None):
        """log of the pdf / pmf as scipy.stats evaluates it (distributions.py:110-131, 225-247, 323-352, 428-449, 531-552)"""
        y, mu = np.asarray(y, dtype=np.float64), np.asarray(mu, dtype=np.float64)
        w = np.ones_like(mu) if weights is None else np.asarray(weights, dtype=np.float64)
        if self.name == "normal":
            return scipy.stats.norm.logpdf(y, loc=mu, scale=self.scale / w)
        if self.name == "binomial":
            return scipy.stats.binom.logpmf(y, self.levels, mu / self.levels)
        if self.name == "poisson":
            return scipy.stats.poisson.logpmf(y, mu=mu * w)
        if self.name == "gamma":
            nu = w / self.scale
            return scipy.stats.gamma.logpdf(x=y, a=nu, scale=mu / nu)
        return scipy.stats.invgauss.logpdf(y, mu, scale=self.scale / w)

    def phi(self, y, mu, edof, weights):
        """the scale: known, or the Pearson estimate (distributions.py:53-78)"""
        if self._known_scale:
            return self.scale ** 2
        y, mu = np.asarray(y, dtype=np.float64), np.asarray(mu, dtype=np.float64)
        return np.sum(np.asarray(weights) * self.V(mu) ** -1 * (y - mu) ** 2) / (len(mu) - edof)

    def sample(self, mu):
        """random draws with expectation mu (distributions.py:187-205, 293-308, 399-412, 496-515, 599-612): host RNG"""
        name, scale = self.name, self.scale
        if name == "normal":
            return np.random.normal(loc=mu, scale=scale if scale else 1.0, size=None)
        if name == "binomial":
            return np.random.binomial(n=self.levels, p=mu / self.levels, size=None)
        if name == "poisson":
            return np.random.poisson(lam=mu, size=None)
        if name == "gamma":
            shape = 1.0 / scale
            return np.random.gamma(shape=shape, scale=mu / shape, size=None)
        return np.random.wald(mean=mu, scale=scale, size=None)


class Link:
    def __init__(self, name):
        self.name = name

    def __repr__(self):  # as an unfitted reference GAM shows its link argument
        return repr(self.name)

    def link(self, mu, dist):
        """g(mu) (links.py:32-46, 91-105, 151-165, 210-224, 269-283)"""
        mu = np.asarray(mu, dtype=np.float64)
        with np.errstate(all="ignore"):
            return {"identity": lambda: mu, "logit": lambda: np.log(mu) - np.log(dist.levels - mu), "log": lambda: np.log(mu),
                    "inverse": lambda: mu ** -1.0}.get(self.name, lambda: mu ** -2.0)()

    def mu(self, lp, dist):
        """g^-1(lp) (links.py:48-62, 107-122, 167-181, 226-240, 285-299)"""
        lp = np.asarray(lp, dtype=np.float64)
        with np.errstate(all="ignore"):
            if self.name == "identity":
                return lp
            if self.name == "logit":
                e = np.exp(lp)
                return float(dist.levels) * e / (e + 1.0)
            if self.name == "log":
                return np.exp(lp)
            if self.name == "inverse":
                return lp ** -1.0
            return lp ** -0.5

    def gradient(self, mu, dist):
        """g'(mu) (links.py:64-77, 124-137, 183-196, 242-255, 301-314)"""
        mu = np.asarray(mu, dtype=np.float64)
        with np.errstate(all="ignore"):
            return {"identity": lambda: np.ones_like(mu), "logit": lambda: dist.levels / (mu * (dist.levels - mu)),
                    "log": lambda: 1.0 / mu, "inverse": lambda: -1 * mu ** -2.0}.get(self.name, lambda: -2 * mu ** -3.0)()


DISTRIBUTIONS = ("normal", "poisson", "binomial", "gamma", "inv_gauss")
LINKS = ("identity", "log", "logit", "inverse", "inv_squared")
CALLBACKS = ("deviance", "diffs", "accuracy", "coef")
_DEVICE_KEYS = ("_engine", "_engine_key", "_gram_cache", "_dev_memo")
PLURAL = ["feature", "dtype", "lam", "n_splines", "spline_order", "constraints", "penalties", "basis",
          "edge_knots_"]


class _LazyStats(dict):
    """statistics_ with entries that are computed on first access: 'edof_per_coef' (an m x m factorisation on the
    host, only summary() reads it) and, for grid-search candidates, 'p_values' (a pseudo-inverse per term,
    pygam.py:1241-1295, read for the best model and by summary())."""

    def __init__(self, owner, *a, lazy=("edof_per_coef",), **kw):
        super().__init__(*a, **kw)
        self._owner = owner
        self._lazy = set(lazy)
        self._ctx = None  # (G, E^T E) on the host, for edof_per_coef

    def __missing__(self, key):
        if key in self._lazy:
            if key == "p_values":
                self[key] = self._owner._estimate_p_values()
            else:
                self[key] = self._owner._estimate_edof_per_coef(self._ctx)
                self._ctx = None
            return self[key]
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._lazy

    def get(self, key, default=None):
        return self[key] if key in self else default

    def materialise(self):
        for key in list(self._lazy):
            self[key]
        return dict(self)

    def __reduce__(self):  # pickles / deep-copies as a plain dict with everything computed
        return (dict, (self.materialise(),))


def _sig_code(p_value):
    """utils.sig_code (utils.py:569-590)"""
    assert 0 <= p_value <= 1, "p_value must be on [0, 1]"
    for cut, code in ((0.001, "***"), (0.01, "**"), (0.05, "*"), (0.1, ".")):
        if p_value < cut:
            return code
    return " "


def _flatten(v):
    out = []
    for x in v:
        if isinstance(x, (list, tuple, np.ndarray)):
            out.extend(_flatten(x))
        else:
            out.append(x)
    return out


def _combine(*grids):
    """cartesian product in the order of utils.combine (utils.py:792-816)"""
    out = [[]]
    for g in grids:
        out = [prev + [v] for prev in out for v in g]
    return out


class GAM:
    def __init__(self, terms="auto", max_iter=100, tol=1e-4, distribution="normal", link="identity",
                 callbacks=("deviance", "diffs"), fit_intercept=True, verbose=False, **kwargs):
        self.max_iter = max_iter
        self.tol = tol
        self.distribution = distribution
        self.link = link
        self.callbacks = list(callbacks)
        self.verbose = verbose
        self.terms = TermList(terms) if isinstance(terms, Term) else terms
        self.fit_intercept = fit_intercept
        for k, v in kwargs.items():
            if k not in PLURAL:
                raise TypeError(f"__init__() got an unexpected keyword argument {k}")
            setattr(self, k, v)
        self._pending_plural = {k: v for k, v in kwargs.items()}
        # pygam.py:187-189
        self._constraint_lam = 1e9
        self._constraint_l2 = 1e-3
        self._constraint_l2_max = 1e-1
        self._expectile = None
        self._engine = None
        self._engine_key = None
        self._gram_cache = None
        self._dev_memo = None
        self._validate_params(check_terms=False)  # the reference validates at fit time (pygam.py:881); see there

    # ---- parameters ----
    _param_names = ("max_iter", "tol", "callbacks", "fit_intercept", "verbose", "terms")

    def __repr__(self):
        """Core.__repr__ (core.py:128-138): the class name and the user-facing parameters, wrapped at 70 columns"""
        from .utils import nice_repr

        return nice_repr(self.__class__.__name__, self.get_params(), line_width=70, line_offset=3, decimals=4)

    def get_params(self, deep=False):
        names = list(self._param_names)
        if type(self) is GAM:
            names += ["distribution", "link"]
        out = {k: getattr(self, k) for k in names}
        for extra in ("scale", "expectile"):
            if hasattr(self, extra):
                out[extra] = getattr(self, extra)
        out.update(self._pending_plural)  # lam=, n_splines=, ... given to the constructor, until a fit moves them to the terms
        if deep:
            for k in ("coef_", "statistics_", "logs_"):
                if hasattr(self, k):
                    out[k] = getattr(self, k)
        return out

    def set_params(self, deep=False, force=False, **parameters):
        for k, v in parameters.items():
            if k == "lam" or (k in PLURAL and isinstance(self.terms, TermList) and k != "lam"):
                self._set_plural(k, v)
            elif k in self.get_params(deep=deep) or force or k in PLURAL:
                setattr(self, k, v)
            elif self.verbose:
                warnings.warn(f"parameter {k} not found")
        return self

    def _set_plural(self, name, value):
        if isinstance(self.terms, TermList):
            self.terms._set_plural(name, value)
        else:
            self._pending_plural[name] = value
            setattr(self, name, value)

    @property
    def lam(self):
        if isinstance(self.terms, TermList):
            return self.terms.lam
        return self._pending_plural.get("lam", 0.6)

    @lam.setter
    def lam(self, value):
        if isinstance(self.__dict__.get("terms"), TermList):
            self.terms.lam = value
        else:
            self.__dict__.setdefault("_pending_plural", {})["lam"] = value

    def __getattr__(self, name):
        """gam.n_splines, gam.spline_order, ...: the values of every term but the intercept (MetaTermMixin.__getattr__,
        terms.py:482-495).  Only reached when normal lookup fails."""
        if name in PLURAL and not name.startswith("_"):
            terms = self.__dict__.get("terms")
            if isinstance(terms, TermList):
                return [getattr(t, name, None) for t in terms if not t.isintercept]
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    @property
    def _is_fitted(self):
        return hasattr(self, "coef_")

    def _validate_params(self, check_terms=True):
        """pygam.py:225-284.  The constructor runs it without the check of `terms` (distribution and link become
        objects right away); a wrong `terms` raises at fit time, where the reference validates
        (pygam/tests/test_GAM_params.py:49-56)."""
        if not (isinstance(self.fit_intercept, bool)):
            raise ValueError(f"fit_intercept must be type bool, but found {self.fit_intercept.__class__}")
        if check_terms and not (isinstance(self.terms, (TermList, Term, type(None))) or
                                (isinstance(self.terms, str) and self.terms == "auto")):
            raise ValueError(f"terms must be a TermList, but found terms = {self.terms}")
        if not (isinstance(self.max_iter, (int, np.integer)) and self.max_iter >= 1):
            raise ValueError(f"max_iter must be int >= 1, but found max_iter = {self.max_iter}")
        if not (np.isfinite(self.tol) and self.tol > 0):
            raise ValueError(f"tol must be float > 0, but found tol = {self.tol}")
        if isinstance(self.distribution, str):
            if self.distribution not in DISTRIBUTIONS:
                raise ValueError(f"unsupported distribution {self.distribution}")
            self.distribution = Distribution(self.distribution, scale=getattr(self, "scale", None),
                                             levels=getattr(self, "levels", 1))
        elif not isinstance(self.distribution, Distribution):
            raise NotImplementedError("custom Distribution objects have no CUDA implementation")
        if isinstance(self.link, str):
            if self.link not in LINKS:
                raise ValueError(f"unsupported link {self.link}")
            self.link = Link(self.link)
        elif not isinstance(self.link, Link):
            raise NotImplementedError("custom Link objects have no CUDA implementation")
        # Deviance() / Diffs() / Accuracy() / Coef() objects (pygam_b200.callbacks) are the built-ins by another name
        self.callbacks = [builtin_name(c) or c for c in self.callbacks]
        for c in self.callbacks:
            if not (isinstance(c, str) and c in CALLBACKS):
                raise NotImplementedError(
                    "only the built-in callbacks 'deviance', 'diffs', 'accuracy', 'coef' are supported: "
                    "arbitrary CallBack objects receive n-sized y/mu arrays (pygam.py:780), which this "
                    "engine never materialises on the host")

    def _validate_data_dep_params(self, data, col_minmax):
        """pygam.py:286-331: default terms, the intercept, plural kwargs, term compilation"""
        F = data.n_features
        if isinstance(self.terms, str) and self.terms == "auto":
            self.terms = TermList(*[SplineTerm(feat, verbose=self.verbose) for feat in range(F)])
        elif self.terms is None:
            self.terms = TermList()
        else:
            self.terms = TermList(self.terms, verbose=self.verbose)
        if self.fit_intercept:
            self.terms = self.terms + Intercept()
        if len(self.terms) == 0:
            raise ValueError("At least 1 term must be specified")
        for k, v in list(self._pending_plural.items()):
            self.terms._set_plural(k, v)
            if k in self.__dict__:
                del self.__dict__[k]
        self._pending_plural = {}
        n_levels = {}
        for t in self.terms:
            if isinstance(t, FactorTerm):
                lo, hi = col_minmax[t.feature]
                n_levels[t.feature] = count_levels(data, t.feature, lo, hi)
        self.terms.compile(_ShapeOnly(data.n, F), col_minmax=col_minmax, n_levels=n_levels)
        feats = self.terms.features_used()
        if feats and max(feats) >= F:
            raise ValueError(f"terms require feature {max(feats)}, but X has only {F} dimensions")

    # ---- data ingestion ----
    def _prepare_data(self, X, y=None, weights=None, fitting=True):
        """host arrays -> DeviceData, with the reference's input checks (utils.py:118-355)"""
        if isinstance(X, DeviceData):
            data = X
            if y is not None or weights is not None:
                data = DeviceData(X.Xt, y if y is not None else X.y, weights if weights is not None else X.w)
        else:
            Xh = np.asarray(X)
            if Xh.dtype.kind not in "fiub":
                try:
                    Xh = Xh.astype("float")
                except ValueError as e:
                    raise ValueError("X data must be type int or float, but found type: " + str(Xh.dtype)) from e
            if Xh.ndim == 0:  # a single value for a one-feature model (make_2d, utils.py:263-290)
                Xh = Xh.reshape(1, 1)
            if Xh.ndim == 1:
                Xh = Xh[:, None]
            if Xh.ndim != 2:
                raise ValueError(f"X data must have 2 dimensions, but found {Xh.ndim}")
            yh = None
            if y is not None:
                yh = np.asarray(y)
                if yh.dtype.kind not in "fiub":
                    raise ValueError("y data must be type int or float, but found type: " + str(yh.dtype))
                yh = np.asarray(yh, dtype=np.float64).ravel()
                if len(yh) != Xh.shape[0]:
                    raise ValueError(f"Inconsistent input and output data shapes. found X: {Xh.shape} "
                                     f"and y: {yh.shape}")
            wh = None
            if weights is not None:
                wh = np.array(weights).astype("f").ravel()  # float32 cast, pygam.py:889
                if not np.isfinite(wh).all():
                    raise ValueError("sample weights data must not contain Inf nor NaN")
                if len(wh) != Xh.shape[0]:
                    raise ValueError(f"Inconsistent input data shapes. found {Xh.shape[0]} and {len(wh)}")
            data = DeviceData.from_host(Xh, yh, wh)
        if data.n < 1:
            raise ValueError("X data should have at least 1 samples")
        if fitting and data.y is not None:
            self._check_y(data.y)
        return data

    def _check_y(self, y):
        """utils.check_y (utils.py:204-248): finite and inside the link's domain"""
        lo, hi, nonfinite = vector_range(y)
        if nonfinite:
            raise ValueError("y data must not contain Inf nor NaN")
        lk, levels = self.link.name, float(self.distribution.levels)
        bad = False
        if lk == "logit":
            bad = lo < 0 or hi > levels
        elif lk == "log":
            bad = lo < 0
        if bad:
            raise ValueError(f"y data is not in domain of {lk} link function. "
                             f"Expected domain: {self._link_domain()}, but found "
                             f"{[lo, hi]}")

    def _link_domain(self):
        return {"logit": [0.0, float(self.distribution.levels)], "log": [0.0, np.inf]}.get(
            self.link.name, [-np.inf, np.inf])

    # ---- engine ----
    def _get_engine(self, data):
        # the analytic null vectors are part of the engine: they can change between fits of the same descriptor
        # (an order-0 basis whose kept edge knots no longer span the data, terms.null_vectors)
        desc_key = (descriptor_key(self.terms, data.n_features), self.distribution.name, self.link.name,
                    float(self.distribution.levels), self._expectile, data.n_features, str(data.device),
                    self.terms.null_vectors().tobytes())
        if self._engine is None or self._engine_key != desc_key:
            self._engine = get_backend().make_engine(self.terms, self.distribution.name, self.link.name,
                                                     self.distribution.levels, self._expectile,
                                                     data.n_features, data.device)
            self._engine_key = desc_key
            self._gram_cache = None
            self._dev_memo = None
        return self._engine

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in _DEVICE_KEYS:
                new.__dict__[k] = v  # device state is shared, never copied
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def _clone_for_candidate(self):
        """what gridsearch's deepcopy(self) (pygam.py:2045) amounts to for a candidate that only changes lam:
        terms, distribution and logs are copied, the device state is shared, fitted statistics are not carried"""
        new = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            if k in _DEVICE_KEYS or k == "link":
                new.__dict__[k] = v
            elif k == "statistics_":
                continue
            elif k in ("terms", "distribution", "callbacks", "_pending_plural", "logs_", "coef_"):
                new.__dict__[k] = copy.deepcopy(v)
            else:
                new.__dict__[k] = v
        return new

    def __getstate__(self):
        d = dict(self.__dict__)
        for k in _DEVICE_KEYS:
            d[k] = None
        return d

    # ---- the solver ----
    def _cholesky_check(self, eng, A):
        """GAM._cholesky (pygam.py:499-537): retry with 10x the diagonal loading, literally"""
        A = np.array(A, dtype=np.float64)
        eye = np.eye(A.shape[0])
        l2 = self._constraint_l2
        while l2 <= self._constraint_l2_max:
            if eng.chol_ok(A):
                self._constraint_l2 = l2
                return A
            if self.verbose:
                warnings.warn("Matrix is not positive definite. \nIncreasing l2 reg by factor of 10.", stacklevel=2)
            A -= l2 * eye
            l2 *= 10
            A += l2 * eye
        raise NotPositiveDefiniteError("Matrix is not positive \ndefinite.")

    def _initial_estimate(self, eng, data):
        """pygam.py:665-710: sqrt(EPS) for LinearGAM, else solve(X^T X + sqrt(EPS) I, X^T link(y'))"""
        m = self.terms.n_coefs
        if isinstance(self, LinearGAM):
            return np.ones(m) * np.sqrt(EPS)
        eng.gram_pass(data, None, mode=_lib.MODE_INIT)
        sc = eng.gram_scalars()
        assert sc[_lib.GS_NONFINITE] == 0, "transformed response values should be well-behaved."
        eng.set_penalty(np.eye(m) * np.sqrt(EPS))
        sol = eng.solve()
        if sol["info"] != 0:
            raise np.linalg.LinAlgError("Singular matrix")
        return sol["coef"]

    def _coef_independent_gram(self):
        """Normal/identity without asymmetric weights: W and z do not depend on the coefficients,
        so G and r are computed once per data set"""
        return (self.distribution.name == "normal" and self.link.name == "identity" and self._expectile is None)

    def _pirls(self, data):
        """One penalised IRLS fit (pygam.py:715-822) on device-resident data"""
        eng = self._get_engine(data)
        m = self.terms.n_coefs
        n = self._n_global
        if (not self._is_fitted or len(self.coef_) != m or not np.isfinite(self.coef_).all()):
            self.coef_ = self._initial_estimate(eng, data)
        assert np.isfinite(self.coef_).all(), f"coefficients should be well-behaved, but found: {self.coef_}"
        self.coef_ = np.asarray(self.coef_, dtype=np.float64)

        Pn, Pr = self.terms._penalties(split=True)
        S = np.eye(m) * np.sqrt(EPS)
        constrained = self.terms.hasconstraint
        if not constrained:
            EtE = self._cholesky_check(eng, S + (Pn + Pr))
            eng.set_penalty(EtE - Pn, Pn)

        cache_ok = self._coef_independent_gram() and not constrained
        cache = getattr(self, "_gram_cache", None) if cache_ok else None
        if cache is not None and cache.get("data") is not data:
            cache = None
        diff = np.inf
        for _ in range(self.max_iter):
            if constrained:
                Pn, Pr = self.terms._penalties(split=True)
                Cn, Cr = self.terms._constraints(self.coef_, self._constraint_lam, self._constraint_l2,
                                                      split=True)
                # the reference factors S + P + C (pygam.py:758-761); any diagonal loading its
                # retry loop adds lands in the part that is not null-annihilating
                EtE = self._cholesky_check(eng, S + (Pn + Pr) + (Cn + Cr))
                eng.set_penalty(EtE - (Pn + Cn), Pn + Cn)

            if cache is not None:
                # G, r cached on the device; only the log scalars need the data (O(k) pass)
                eng.out.copy_(cache["out"])
                dev, ncorrect = self._dev_at(eng, data, self.coef_)
                nused = n
            else:
                eng.gram_pass(data, self.coef_)
                sc = eng.gram_scalars()
                nused, dev, ncorrect = sc[_lib.GS_NUSED], sc[_lib.GS_DEV], sc[_lib.GS_NCORRECT]
                if nused == 0:
                    raise OptimizationError(
                        "PIRLS optimization has diverged.\n"
                        "Try increasing regularization, or specifying an initial value for self.coef_")
                if cache_ok and nused == n:
                    cache = dict(data=data, out=eng.out.clone())
                    self._gram_cache = cache
            self._log_start(dev, ncorrect, nused)

            # fewer kept rows than coefficients: the reference's SVD is cropped to that many singular triplets
            # (pygam.py:789-799); the statistics below use the rank of the last iteration like the reference's U1
            self._solve_rank = int(min(m, n, nused))
            if self._solve_rank < m:
                sol = eng.solve_truncated(self._solve_rank, coef_prev=self.coef_)
            else:
                sol = eng.solve(coef_prev=self.coef_)
            if sol["info"] == 2:
                raise ValueError("QR decomposition produced NaN or Inf. Check X data.")
            if sol["info"] != 0:
                raise NotPositiveDefiniteError("penalised normal equations are not positive definite")
            coef_new = sol["coef"]
            diff = sol["diff"]
            self.coef_ = coef_new
            self._log_end(diff)
            if diff < self.tol:
                break

        self._estimate_model_statistics(eng, data)
        if diff >= self.tol:
            print("did not converge")  # pygam.py:818-822
        return

    def _dev_at(self, eng, data, coef):
        """unweighted deviance / accuracy count at `coef` through the O(k) statistics pass"""
        memo = getattr(self, "_dev_memo", None)
        key = np.asarray(coef, dtype=np.float64).tobytes()
        if memo is not None and memo[0] is data and memo[1] == key:
            return memo[2], memo[3]
        sc, _, _ = eng.stats_pass(data, coef)
        self._dev_memo = (data, key, sc[_lib.SS_DEV], sc[_lib.SS_NCORRECT])
        return sc[_lib.SS_DEV], sc[_lib.SS_NCORRECT]

    # ---- logs (callbacks.py:110-233, pygam.py:824-858) ----
    def _log_start(self, dev, ncorrect, nused):
        if "deviance" in self.callbacks:
            self.logs_["deviance"].append(float(dev))
        if "accuracy" in self.callbacks:
            self.logs_["accuracy"].append(float(ncorrect) / float(nused))
        if "coef" in self.callbacks:
            self.logs_["coef"].append(np.array(self.coef_))

    def _log_end(self, diff):
        if "diffs" in self.callbacks:
            self.logs_["diffs"].append(float(diff))

    # ---- statistics (pygam.py:993-1228) ----
    def _estimate_model_statistics(self, eng, data):
        n = self._n_global
        dist = self.distribution
        st = self.statistics_
        # edof and cov from the LAST iteration's Gram matrix (weights at the previous
        # coefficients), as the reference's B and U1 are (pygam.py:815-817)
        if getattr(self, "_solve_rank", self.terms.n_coefs) < self.terms.n_coefs:
            sol = eng.solve_truncated(self._solve_rank, want_cov=True)
        else:
            sol = eng.solve(coef_prev=None, want_edof=True, want_cov=True)
        if isinstance(st, _LazyStats):  # edof_per_coef on first access, from the last iteration's G and penalty
            mm = self.terms.n_coefs
            st._ctx = (eng.out[:mm * mm].view(mm, mm).cpu().numpy().copy(), eng.penalty_host())
        st["edof"] = sol["edof"]
        sc, _, _ = eng.stats_pass(data, self.coef_)
        self._dev_memo = (data, np.asarray(self.coef_, dtype=np.float64).tobytes(), sc[_lib.SS_DEV],
                          sc[_lib.SS_NCORRECT])
        if not dist._known_scale:
            dist.scale = float(np.sqrt(sc[_lib.SS_PEARSON] / (n - st["edof"])))  # distributions.py:78
        st["scale"] = dist.scale
        st["cov"] = sol["cov"] * dist.scale ** 2
        # cov = A^-1 G A^-1 is positive semi-definite; clamp the -1e-20 rounding of a zero variance
        st["se"] = np.maximum(st["cov"].diagonal(), 0.0) ** 0.5
        # second O(k) pass: log-likelihood and the null model need scale and mean(y)
        ybar = sc[_lib.SS_SUMY] / n
        sc2, _, _ = eng.stats_pass(data, self.coef_, scale=dist.scale, ybar=ybar,
                                   poisson_round=isinstance(self, PoissonGAM))
        ll = sc2[_lib.SS_LOGLIK]
        st["AIC"] = -2 * ll + 2 * st["edof"] + 2 * (not dist._known_scale)
        st["AICc"] = st["AIC"] + 2 * (st["edof"] + 1) * (st["edof"] + 2) / (n - st["edof"] - 2)
        dev = sc[_lib.SS_WDEV]
        null_dev = sc2[_lib.SS_NULL_WDEV]
        null_ll = sc2[_lib.SS_NULL_LOGLIK]
        st["pseudo_r2"] = OrderedDict([("explained_deviance", 1.0 - dev / null_dev),
                                       ("McFadden", ll / null_ll),
                                       ("McFadden_adj", 1.0 - (ll - st["edof"]) / null_ll)])
        st["GCV"], st["UBRE"] = self._estimate_GCV_UBRE(dev, n, st["edof"])
        st["loglikelihood"] = ll
        div = dist.scale ** 2 if dist.name == "normal" else dist.scale  # scaled deviance, pygam.py:1054-1056
        st["deviance"] = dev / div
        st["p_values"] = self._estimate_p_values()

    def _estimate_GCV_UBRE(self, dev, n, edof, gamma=1.4, add_scale=True):
        """pygam.py:1156-1228; `~add_scale` of the reference is -2 for add_scale=True"""
        if gamma < 1:
            raise ValueError(f"gamma scaling should be greater than 1, but found gamma = {gamma}")
        GCV = UBRE = None
        if self.distribution._known_scale:
            scale = self.distribution.scale
            UBRE = 1.0 / n * dev - (-2 if add_scale else -1) * scale + 2.0 * gamma / n * edof * scale
        else:
            GCV = (n * dev) / (n - gamma * edof) ** 2
        return GCV, UBRE

    def _estimate_p_values(self):
        return [self._compute_p_value(i) for i in range(len(self.terms))]

    def _compute_p_value(self, term_i):
        """pygam.py:1241-1295 (Wood 2006, 4.8.5): m-sized host arithmetic on cov and coef"""
        idxs = self.terms.get_coef_indices(term_i)
        cov = self.statistics_["cov"][idxs][:, idxs]
        coef = np.array(self.coef_[idxs])
        if isinstance(self.terms[term_i], SplineTerm):
            coef -= coef.mean()
        # pinv and rank of the block as scipy.linalg.pinv computes them (pygam.py:1283), with one difference.  A block
        # with fewer distinct x than splines (wage: 7 years on 20 splines) has directions of exactly zero variance;
        # in the reference's cov (a product B B^T from its SVD) they sit at 1e-15 of the largest singular value, below
        # scipy's cut-off max(M, N) eps, and its rank is 7.  cov here comes through the normal equations and carries
        # eps * cond(G + P) there (1e-11 .. 1e-10 on wage): above that cut-off, which would make the rank 19 and the
        # p-value 0.40 instead of 0.006.  The cut-off is therefore sqrt(eps) of the largest singular value: the same
        # rank as the reference wherever the true singular values of the block stay above 1.5e-8 of the largest.
        U, sv, Vt = np.linalg.svd(cov)
        rank = int(np.sum(sv > sv[0] * np.sqrt(EPS))) if sv[0] > 0 else 0
        inv_cov = (Vt[:rank].T / sv[:rank]).dot(U[:, :rank].T)
        score = coef.T.dot(inv_cov).dot(coef)
        if self.distribution._known_scale:
            return 1 - scipy.stats.chi2.cdf(x=score, df=rank)
        score = score / rank
        return 1 - scipy.stats.f.cdf(score, rank, self.statistics_["n_samples"] - self.statistics_["edof"])

    def _estimate_edof_per_coef(self, ctx):
        """diag(U1 U1^T) of the reference (pygam.py:1033) = diag(R (G + E^T E)^-1 R^T) with R^T R = G, R upper
        triangular (the R of its QR of W.B).  G is singular along the analytic null space of the design and wherever
        a block has fewer distinct x than splines; there the reference's R rows are rounding noise and its split of
        edof over the coefficients is not reproducible (its SUM is).  Here R is the Cholesky factor in the limit of
        exact arithmetic: a dependent column gets a zero row.  Agrees with the reference to 1e-12 on designs without
        null directions, to its own noise (~1e-2 per coefficient) otherwise."""
        if ctx is None:
            return None
        G, EtE = ctx
        m = len(G)
        k = getattr(self, "_solve_rank", m)
        if k < m:
            # rows < coefficients: U1 is the k x k corner of U (pygam.py:797), so diag(U1 U1^T) has k entries, one per
            # ROW of the k x m factor R of the QR -- fewer values than coefficients, which is why summary() prints no
            # per-term edof then (pygam/tests/test_GAM_methods.py:93-107).  U1 = R V_k D_k^-1 over the k leading
            # eigenpairs of A = G + E^T E, hence diag(R A_k^+ R^T); its sum is the edof of the cropped solve.
            R = np.zeros((k, m))
            d, row = np.diag(G), 0
            for j in range(m):
                if row == k:
                    break
                piv = G[j, j] - R[:row, j] @ R[:row, j]
                if not piv > 1e-12 * d[j]:
                    continue
                R[row, j:] = (G[j, j:] - R[:row, j] @ R[:row, j:]) / np.sqrt(piv)
                row += 1
            lam_, V = np.linalg.eigh(G + EtE)
            Vk, lk = V[:, ::-1][:, :k], lam_[::-1][:k]
            RV = R @ Vk
            return np.einsum("ij,ij->i", RV / lk, RV)
        R = np.zeros((m, m))
        d = np.diag(G)
        for j in range(m):
            piv = G[j, j] - R[:j, j] @ R[:j, j]
            if not piv > 1e-12 * d[j]:
                continue
            R[j, j:] = (G[j, j:] - R[:j, j] @ R[:j, j:]) / np.sqrt(piv)
        Z = np.linalg.solve(G + EtE, R.T)
        return np.einsum("ij,ji->i", R, Z)

    # ---- reporting (pygam.py:917-991, 1645-1806, 2087-2368): small-n host code over coef_, cov and the O(k) passes ----
    def summary(self):
        """pygam.py:1645-1806: the same table"""
        if not self._is_fitted:
            raise AttributeError("GAM has not been fitted. Call fit first.")
        st = self.statistics_
        wd, wr = 47, 58

        def space_row(left, right, total):
            left, right = str(left), str(right)
            return left + " " * (total - len(left) - len(right)) + right

        objective = "UBRE" if self.distribution._known_scale else "GCV"
        dist_name = {"normal": "NormalDist", "binomial": "BinomialDist", "poisson": "PoissonDist", "gamma": "GammaDist",
                     "inv_gauss": "InvGaussDist"}[self.distribution.name]
        link_name = {"identity": "IdentityLink", "logit": "LogitLink", "log": "LogLink", "inverse": "InverseLink",
                     "inv_squared": "InvSquaredLink"}[self.link.name]
        rows = [(space_row("Distribution:", dist_name, wd), space_row("Effective DoF:", np.round(st["edof"], 4), wr)),
                (space_row("Link Function:", link_name, wd), space_row("Log Likelihood:", np.round(st["loglikelihood"], 4), wr)),
                (space_row("Number of Samples:", st["n_samples"], wd), space_row("AIC: ", np.round(st["AIC"], 4), wr)),
                ("", space_row("AICc: ", np.round(st["AICc"], 4), wr)),
                ("", space_row(objective + ":", np.round(st[objective], 4), wr)),
                ("", space_row("Scale:", np.round(st["scale"], 4), wr)),
                ("", space_row("Pseudo R-Squared:", np.round(st["pseudo_r2"]["explained_deviance"], 4), wr))]
        out = ["{:47} {:58}".format(self.__class__.__name__, ""), "{:47} {:58}".format("=" * wd, "=" * wr)]
        out += ["{:47} {:58}".format(a[:wd], b[:wr]) for a, b in rows]
        out.append("=" * 106)
        fmt = [("Feature Function", 33), ("Lambda", 20), ("Rank", 12), ("EDoF", 12), ("P > x", 12), ("Sig. Code", 12)]
        line = " ".join("{:%d}" % w for _, w in fmt)
        out.append(line.format(*[h for h, _ in fmt]))
        out.append(line.format(*["=" * w for _, w in fmt]))
        epc = st["edof_per_coef"]
        for i, term in enumerate(self.terms):
            if epc is not None and len(epc) == len(self.coef_):
                edof = np.round(epc[self.terms.get_coef_indices(i)].sum(), 1)
            else:
                edof = ""
            pv = st["p_values"][i]
            cells = [repr(term), "" if term.isintercept else np.round(_flatten(term.lam), 4), f"{term.n_coefs}", f"{edof}",
                     "%.2e" % pv, _sig_code(pv)]
            out.append(line.format(*[str(c)[:w] for c, (_, w) in zip(cells, fmt)]))
        out.append("=" * 106)
        out.append("Significance codes:  0 '***' 0.001 '**' 0.01 '*' 0.05 '.' 0.1 ' ' 1")
        out.append("")
        out.append("WARNING: Fitting splines and a linear function to a feature introduces a model identifiability problem\n"
                   "         which can cause p-values to appear significant when they are not.")
        out.append("")
        out.append("WARNING: p-values calculated in this manner behave correctly for un-penalized models or models with\n"
                   "         known smoothing parameters, but when smoothing parameters have been estimated, the p-values\n"
                   "         are typically lower than they should be, meaning that the tests reject the null too readily.")
        print("\n".join(out))
        warnings.warn("KNOWN BUG: p-values computed in this summary are likely much smaller than they should be. \n \n"
                      "Please do not make inferences based on these values! \n\n"
                      "Collaborate on a solution, and stay up to date at: \ngithub.com/dswah/pyGAM/issues/163 \n",
                      stacklevel=2)

    def _deviance_rows(self, y, mu, weights=None, scaled=True):
        """Distribution.deviance on host arrays"""
        return self.distribution.deviance(y, mu, scaled=scaled, weights=weights)

    def deviance_residuals(self, X, y, weights=None, scaled=False):
        """pygam.py:940-991: sign(y - mu) sqrt(deviance row); mu on the device, the n-sized result on the host"""
        if not self._is_fitted:
            raise AttributeError("GAM has not been fitted. Call fit first.")
        data = self._prepare_data(X, y, weights, fitting=True)
        yh = data.y.cpu().numpy()
        wh = np.ones_like(yh) if data.w is None else data.w.cpu().numpy().astype(np.float64)
        mu = self.predict_mu(X)
        return np.sign(yh - mu) * self._deviance_rows(yh, mu, weights=wh, scaled=scaled) ** 0.5

    def _estimate_r2(self, X, y=None, weights=None):
        """pygam.py:1114-1154 through the O(k) statistics pass"""
        data = self._prepare_data(X, y, weights, fitting=True)
        eng = self._get_engine(data)
        sc, _, _ = eng.stats_pass(data, self.coef_)
        n = sc[_lib.SS_N]
        sc2, _, _ = eng.stats_pass(data, self.coef_, scale=self.distribution.scale, ybar=sc[_lib.SS_SUMY] / n,
                                   poisson_round=isinstance(self, PoissonGAM))
        ll, null_ll = sc2[_lib.SS_LOGLIK], sc2[_lib.SS_NULL_LOGLIK]
        return OrderedDict([("explained_deviance", 1.0 - sc[_lib.SS_WDEV] / sc2[_lib.SS_NULL_WDEV]),
                            ("McFadden", ll / null_ll), ("McFadden_adj", 1.0 - (ll - self.statistics_["edof"]) / null_ll)])

    def score(self, X, y, weights=None):
        """explained deviance on (X, y) (pygam.py:917-938)"""
        if not self._is_fitted:
            raise AttributeError("GAM has not been fitted. Call fit first.")
        data = self._prepare_data(X, y, weights, fitting=True)
        self._check_predict_X(data)
        return self._estimate_r2(data)["explained_deviance"]

    def _dist_sample(self, mu):
        return self.distribution.sample(mu)

    def sample(self, X, y, quantity="y", sample_at_X=None, weights=None, n_draws=100, n_bootstraps=5, objective="auto"):
        """Simulate from the posterior of the coefficients and smoothing parameters (pygam.py:2087-2210): bootstrap
        refits (each a grid search + a fit on the device), draws from N(coef_b, cov_b) on the host"""
        if quantity not in {"mu", "coef", "y"}:
            raise ValueError(f"`quantity` must be one of 'mu', 'coef', 'y'; got {quantity}")
        coef_draws = self._sample_coef(X, y, weights=weights, n_draws=n_draws, n_bootstraps=n_bootstraps,
                                       objective=objective)
        if quantity == "coef":
            return coef_draws
        if sample_at_X is None:
            sample_at_X = X
        lp = self._modelmat(sample_at_X).dot(coef_draws.T)
        mu = self._link_mu(lp).T
        if quantity == "mu":
            return mu
        return self._dist_sample(mu)

    def _sample_coef(self, X, y, weights=None, n_draws=100, n_bootstraps=1, objective="auto"):
        """pygam.py:2212-2268"""
        if not self._is_fitted:
            raise AttributeError("GAM has not been fitted. Call fit first.")
        if n_bootstraps < 1:
            raise ValueError(f"n_bootstraps must be >= 1; got {n_bootstraps}")
        if n_draws < 1:
            raise ValueError(f"n_draws must be >= 1; got {n_draws}")
        coefs, covs = self._bootstrap_samples_of_smoothing(X, y, weights=weights, n_bootstraps=n_bootstraps,
                                                           objective=objective)
        return self._simulate_coef_from_bootstraps(n_draws, coefs, covs)

    def _bootstrap_samples_of_smoothing(self, X, y, weights=None, n_bootstraps=1, objective="auto"):
        """pygam.py:2270-2337 (Wood 2006, p. 198): y* ~ distribution(mu), lam* by grid search on y*, refit on y"""
        def load_diagonal(cov, load=None):  # utils.py:838-866
            n = cov.shape[0]
            return cov + np.eye(n) * (np.sqrt(EPS) if load is None else load)

        mu = self.predict_mu(X)
        coef_bootstraps = [self.coef_]
        cov_bootstraps = [load_diagonal(self.statistics_["cov"])]
        for _ in range(n_bootstraps - 1):
            y_bootstrap = self._dist_sample(mu)
            gam = copy.deepcopy(self)
            lam_grid = np.exp(np.random.randn(11, len(_flatten(self.lam))) * 6 - 3)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                gam.gridsearch(X, y_bootstrap, weights=weights, lam=lam_grid, objective=objective)
            lam = gam.lam
            gam = copy.deepcopy(self)
            gam.lam = lam
            gam.fit(X, y, weights=weights)
            coef_bootstraps.append(gam.coef_)
            cov_bootstraps.append(load_diagonal(gam.statistics_["cov"]))
        return coef_bootstraps, cov_bootstraps

    def _simulate_coef_from_bootstraps(self, n_draws, coef_bootstraps, cov_bootstraps):
        """pygam.py:2339-2368"""
        idx = np.random.choice(np.arange(len(coef_bootstraps)), size=n_draws, replace=True)
        coef_draws = np.empty((n_draws, len(self.coef_)))
        for b in np.unique(idx):
            draws = np.flatnonzero(idx == b)
            coef_draws[draws] = np.random.multivariate_normal(coef_bootstraps[b], cov_bootstraps[b], size=len(draws))
        return coef_draws

    # ---- public API ----
    def fit(self, X, y=None, weights=None):
        """Fit the model (pygam.py:860-915).  X: (n, F) array-like on the host, or a
        ``DeviceData`` row shard already resident in HBM (then y / weights may live inside it)."""
        self._validate_params()
        data = self._prepare_data(X, y, weights, fitting=True)
        if data.y is None:
            raise ValueError("y is required")
        self._fit_data(data)
        return self

    def _fit_data(self, data):
        col_minmax, bad = column_stats(data)
        if bad.sum() > 0:
            raise ValueError("X data must not contain Inf nor NaN")
        nt = torch.tensor([float(data.n)], dtype=torch.float64, device=data.device)
        self._n_global = int(all_reduce_(nt, "sum").item())
        self._validate_data_dep_params(data, col_minmax)
        self.statistics_ = _LazyStats(self)
        self.statistics_["n_samples"] = self._n_global
        self.statistics_["m_features"] = data.n_features
        if not hasattr(self, "logs_"):  # pygam.py:901-902: created once, so the logs of a refit are appended
            self.logs_ = {c: [] for c in self.callbacks}
        for c in self.callbacks:
            self.logs_.setdefault(c, [])
        self._pirls(data)
        return self

    def _linear_predictor(self, X):
        """pygam.py:385-421"""
        data = self._prepare_data(X, fitting=False)
        self._check_predict_X(data)
        eng = self._get_engine(data)
        _, lp, _ = eng.stats_pass(DeviceData(data.Xt), self.coef_, want_lp=True, reduce=False)
        return lp.cpu().numpy()

    def predict_mu(self, X):
        """pygam.py:423-450"""
        if not self._is_fitted:
            raise AttributeError("GAM has not been fitted. Call fit first.")
        data = self._prepare_data(X, fitting=False)
        self._check_predict_X(data)
        eng = self._get_engine(data)
        _, _, mu = eng.stats_pass(DeviceData(data.Xt), self.coef_, want_mu=True, reduce=False)
        return mu.cpu().numpy()

    def predict(self, X):
        """pygam.py:452-467"""
        return self.predict_mu(X)

    def _check_predict_X(self, data):
        F = self.statistics_["m_features"]
        if data.n_features != F:
            raise ValueError(f"X data must have {F} features, but found {data.n_features}")
        with single_rank():  # this rank's rows only (not memoised in `data`: a fit needs the statistics of all ranks):
            col_minmax, bad = get_backend().column_stats(data)  # one pass gives the finite check and the column ranges
        if bad.sum() > 0:
            raise ValueError("X data must not contain Inf nor NaN")
        for t in self.terms:  # categorical range check, utils.py:308-334
            if isinstance(t, FactorTerm):
                lo, hi = float(t.edge_knots_[0]), float(t.edge_knots_[-1])
                xlo, xhi = float(col_minmax[t.feature, 0]), float(col_minmax[t.feature, 1])
                if xlo < lo or xhi > hi:
                    raise ValueError(f"X data is out of domain for categorical feature {t.feature}. "
                                     f"Expected data on [{lo}, {hi}], but found data on "
                                     f"[{xlo}, {xhi}]")

    def _modelmat(self, X, term=-1):
        """dense model matrix of a (small) X, for inspection (pygam.py:469-497)"""
        data = self._prepare_data(X, fitting=False)
        eng = self._get_engine(data)
        mm = eng.model_matrix(data).cpu().numpy()
        if term == -1:
            return mm
        return mm[:, self.terms.get_coef_indices(term)]

    # ---- intervals and partial dependence ----
    def _link_mu(self, lp):
        """Link.mu on host arrays (links.py:32-314), for the small interval outputs"""
        return self.link.mu(lp, self.distribution)

    def _term_index(self, term):
        return term if term >= 0 else len(self.terms) + term

    def _get_quantiles(self, X, width, quantiles, prediction=False, xform=True, term=-1, lp_var=None):
        """pygam.py:1336-1421: lp + q * sqrt(x^T cov x) per quantile; the n-sized part (lp and the
        quadratic form, basis rows on the fly) is pgb_interval_pass"""
        if quantiles is not None:
            quantiles = np.atleast_1d(quantiles)
        else:
            alpha = (1 - width) / 2.0
            quantiles = [alpha, 1 - alpha]
        for quantile in quantiles:
            if (quantile >= 1) or (quantile <= 0):
                raise ValueError(f"quantiles must be in (0, 1), but found {quantiles}")
        if lp_var is None:
            lp_var = self._lp_and_var(X, term)
        lp, var = lp_var
        var = var.copy()
        if prediction:
            var += self.distribution.scale ** 2
        import scipy.stats

        lines = []
        for quantile in quantiles:
            if self.distribution._known_scale:
                q = scipy.stats.norm.ppf(quantile)
            else:
                q = scipy.stats.t.ppf(quantile, df=self.statistics_["n_samples"] - self.statistics_["edof"])
            lines.append(lp + q * var ** 0.5)
        lines = np.vstack(lines).T
        if xform:
            lines = self._link_mu(lines)
        return lines

    def _lp_and_var(self, X, term=-1):
        data = self._prepare_data(X, fitting=False)
        self._check_predict_X(data)
        eng = self._get_engine(data)
        lp, var = eng.interval_pass(DeviceData(data.Xt), self.coef_, self.statistics_["cov"],
                                    -1 if term == -1 else self._term_index(term))
        return lp.cpu().numpy(), var.cpu().numpy()

    def confidence_intervals(self, X, width=0.95, quantiles=None):
        """pygam.py:1297-1334"""
        if not self._is_fitted:
            raise AttributeError("GAM has not been fitted. Call fit first.")
        return self._get_quantiles(X, width, quantiles, prediction=False)

    def _flatten_mesh(self, Xs, term):
        """pygam.py:1423-1436"""
        n = Xs[0].size
        t = self.terms[term]
        sub = list(t) if t.istensor else [t]
        X = np.zeros((n, self.statistics_["m_features"]))
        for term_, x in zip(sub, Xs):
            X[:, term_.feature] = x.ravel()
        return X

    def generate_X_grid(self, term, n=100, meshgrid=False):
        """pygam.py:1438-1518"""
        if not self._is_fitted:
            raise AttributeError("GAM has not been fitted. Call fit first.")
        t = self.terms[term]
        if t.isintercept:
            raise ValueError("cannot create grid for intercept term")
        if t.istensor:
            Xs = [np.linspace(term_.edge_knots_[0], term_.edge_knots_[1], num=n) for term_ in t]
            Xs = np.meshgrid(*Xs, indexing="ij")
            if meshgrid:
                return tuple(Xs)
            return self._flatten_mesh(Xs, term=term)
        if getattr(t, "edge_knots_", None) is not None:
            x = np.linspace(t.edge_knots_[0], t.edge_knots_[1], num=n)
            if meshgrid:
                return (x,)
            X = np.zeros((n, self.statistics_["m_features"]))
            X[:, t.feature] = x
            if getattr(t, "by", None) is not None:
                X[:, t.by] = 1.0
            return X
        raise TypeError(f"Unexpected term type: {t}")

    def partial_dependence(self, term, X=None, width=None, quantiles=None, meshgrid=False):
        """pygam.py:1520-1643"""
        if not self._is_fitted:
            raise AttributeError("GAM has not been fitted. Call fit first.")
        if not isinstance(term, (int, np.integer)) or isinstance(term, bool):
            raise ValueError(f"term must be an integer, but found term: {term}")
        if (term >= len(self.terms)) or (term < -1):
            raise ValueError(f"Term {term} out of range for model with {len(self.terms)} terms")
        if self.terms[term].isintercept:
            raise ValueError("cannot create grid for intercept term")
        if X is None:
            X = self.generate_X_grid(term=term, meshgrid=meshgrid)
        shape = None
        if meshgrid:
            if not isinstance(X, tuple):
                raise ValueError(f"X must be a tuple of grids if `meshgrid=True`, but found X: {X}")
            shape = X[0].shape
            X = self._flatten_mesh(X, term=term)
        compute_quantiles = (width is not None) or (quantiles is not None)
        data = self._prepare_data(X, fitting=False)
        self._check_predict_X(data)
        eng = self._get_engine(data)
        lp, var = eng.interval_pass(DeviceData(data.Xt), self.coef_, self.statistics_["cov"] if compute_quantiles else None,
                                    self._term_index(term))
        pdep = lp.cpu().numpy()
        if len(self.terms) == 1:
            # the only term of a model without intercept: its partial dependence IS the linear predictor, and the
            # reference's equal each other bit for bit (pygam/tests/test_partial_dependence.py:7-16)
            pdep = self._linear_predictor(X)
        out = [pdep]
        if compute_quantiles:
            out.append(self._get_quantiles(X, width=width, quantiles=quantiles, term=term, xform=False,
                                           lp_var=(pdep, var.cpu().numpy())))
        if meshgrid:
            for i, array in enumerate(out):
                shp = shape + (array.shape[-1],) if array.ndim > 1 else shape
                out[i] = np.reshape(array, shp)
        if compute_quantiles:
            return out
        return out[0]

    def _P(self):
        return self.terms._penalties()

    def _C(self):
        return self.terms._constraints(self.coef_, self._constraint_lam, self._constraint_l2)

    def loglikelihood(self, X, y, weights=None):
        """pygam.py:1059-1112: X goes through the checks of any fitted-model input (check_X: features, finite values,
        categorical ranges), y and weights through those of a fit"""
        data = self._prepare_data(X, y, weights, fitting=True)
        self._check_predict_X(data)
        eng = self._get_engine(data)
        sc, _, _ = eng.stats_pass(data, self.coef_)
        ybar = sc[_lib.SS_SUMY] / data.n
        sc2, _, _ = eng.stats_pass(data, self.coef_, scale=self.distribution.scale, ybar=ybar,
                                   poisson_round=isinstance(self, PoissonGAM))
        return sc2[_lib.SS_LOGLIK]

    def gridsearch(self, X, y=None, weights=None, return_scores=False, keep_best=True, objective="auto",
                   progress=False, **param_grids):
        """Grid search over model parameters, best by GCV / UBRE / AIC / AICc
        (pygam.py:1808-2085): sequential refits, each warm-started from the previous
        candidate.  The data is uploaded once; for coefficient-independent weights
        (LinearGAM) the Gram matrix is accumulated once and every candidate costs one m x m
        solve plus O(k) passes."""
        if not self._is_fitted:
            self._validate_params()
        data = self._prepare_data(X, y, weights, fitting=True)
        if data.y is None:
            raise ValueError("y is required")
        if column_stats(data)[1].sum() > 0:  # check_X before any candidate is fitted (pygam.py:1935-1944); memoised for the fits
            raise ValueError("X data must not contain Inf nor NaN")
        if objective not in ["auto", "GCV", "UBRE", "AIC", "AICc"]:
            raise ValueError("objective mut be in ['auto', 'GCV', 'UBRE', 'AIC', 'AICc'], "
                             f"but found objective = {objective}")
        if self.distribution._known_scale:
            if objective == "GCV":
                raise ValueError("GCV should be used for models withunknown scale")
            if objective == "auto":
                objective = "UBRE"
        else:
            if objective == "UBRE":
                raise ValueError("UBRE should be used for models with known scale")
            if objective == "auto":
                objective = "GCV"
        if not self._is_fitted:
            col_minmax, bad = column_stats(data)
            if bad.sum() > 0:
                raise ValueError("X data must not contain Inf nor NaN")
            self._validate_data_dep_params(data, col_minmax)
        if not bool(param_grids):
            param_grids["lam"] = np.logspace(-3, 3, 11)
        admissible = list(self.get_params()) + PLURAL
        params, grids = [], []
        for param, grid in list(param_grids.items()):
            if param not in admissible:
                raise ValueError(f"unknown parameter: {param}")
            if not (np.iterable(grid) and len(grid) > 1):
                raise ValueError(f"{param} grid must either be iterable of iterables, or an iterable of "
                                 f"lengnth > 1, but found {grid}")
            if any(np.iterable(g) and not isinstance(g, str) for g in grid):
                target_len = len(_flatten(getattr(self.terms, param) if param in PLURAL else getattr(self, param)))
                cartesian = not isinstance(grid, np.ndarray) or grid.ndim != 2
                grid = [np.atleast_1d(g) for g in grid]
                msg = f"{param} grid should have {target_len} columns, but found grid with {len(grid)} columns"
                if cartesian:
                    if len(grid) != target_len:
                        raise ValueError(msg)
                    grid = _combine(*grid)
                if not all(len(sub) == target_len for sub in grid):
                    raise ValueError(msg)
            params.append(param)
            grids.append(list(grid))
        param_grid_list = [dict(zip(params, cand)) for cand in _combine(*grids)]

        best_model, best_score = None, np.inf
        scores, models = [], []
        if self._is_fitted:
            models.append(self)
            scores.append(self.statistics_[objective])
            best_model, best_score = models[-1], scores[-1]
        if self._batched_gridsearch_ok(data, params):
            # LinearGAM over a lam grid: one Gram pass, all candidates solved as one batch (split over the ranks)
            done = self._gridsearch_linear_batched(data, param_grid_list, objective, models, scores)
            if done:
                param_grid_list = []
                for gam, score in zip(models, scores):
                    if score < best_score:
                        best_model, best_score = gam, score
        for param_grid in param_grid_list:
            try:
                gam = copy.deepcopy(self)
                gam.set_params(**param_grid)
                if models:
                    gam.coef_ = np.array(models[-1].coef_)
                gam._validate_params()
                gam._fit_data(data)
            except ValueError as error:
                if self.verbose:
                    warnings.warn(str(error) + "\non model with params:\n" + str(param_grid) + "\nskipping...\n")
                continue
            # candidates share the device engine and the cached Gram matrix
            for k in _DEVICE_KEYS:
                self.__dict__[k] = gam.__dict__.get(k)
            models.append(gam)
            scores.append(gam.statistics_[objective])
            if scores[-1] < best_score:
                best_model, best_score = models[-1], scores[-1]
        if len(models) == 0:
            if self.verbose:
                warnings.warn("No models were fitted.")
            return self
        if keep_best:
            for k, v in best_model.__dict__.items():
                if k not in _DEVICE_KEYS:
                    self.__dict__[k] = copy.deepcopy(v)  # a candidate's lazy statistics_ copies as a plain, complete dict
        if return_scores:
            return OrderedDict(zip(models, scores))
        return self


    # ---- grid search of a model whose weights do not depend on the coefficients ----
    def _batched_gridsearch_ok(self, data, params):
        return (self._coef_independent_gram() and list(params) == ["lam"] and data.w is None
                and not self.terms.hasconstraint and set(self.callbacks) <= {"deviance", "diffs", "coef"})

    def _gridsearch_linear_batched(self, data, param_grid_list, objective, models, scores):
        """pygam.py:2042-2070 for Normal / identity without weights or constraints.  G = X^T X and r = X^T y do not
        depend on lam or on the coefficients, so the whole grid costs ONE Gram pass, one O(k) pass for the moments
        of y, and K independent m x m solves (a batch, split over the ranks).  Every quantity the reference
        recomputes from the data per candidate is a quadratic form here: the deviance at beta is
        y^T y - 2 beta^T r + beta^T G beta.  The warm-start chain (pygam.py:2051-2053) only decides how many
        iterations a candidate logs: the first solve from the previous candidate's coefficients already lands on
        the solution, and a second iteration runs iff its relative change was >= tol.
        Returns False (nothing appended) when the quadratic forms would cancel too many digits."""
        eng = self._get_engine(data)
        m = self.terms.n_coefs
        nt = torch.tensor([float(data.n)], dtype=torch.float64, device=data.device)
        n = int(all_reduce_(nt, "sum").item())
        cache = self._gram_cache if (self._gram_cache is not None and self._gram_cache.get("data") is data) else None
        if cache is None:
            eng.gram_pass(data, np.zeros(m))
            if eng.gram_scalars()[_lib.GS_NUSED] != n:
                return False
            cache = dict(data=data, out=eng.out.clone())
            self._gram_cache = cache
        else:
            eng.out.copy_(cache["out"])
        host = eng.out.cpu().numpy()
        G, r = host[:m * m].reshape(m, m).copy(), host[m * m:m * m + m].copy()
        sc0, _, _ = eng.stats_pass(data, np.zeros(m))
        yy, sy = float(sc0[_lib.SS_DEV]), float(sc0[_lib.SS_SUMY])

        def dev_at(beta):
            return yy - 2.0 * float(beta @ r) + float(beta @ (G @ beta))

        # candidates in grid order.  The penalty is linear in lam (terms.py:307-349: sum_k lam_k P_k, block by block), so
        # the K penalties are one tensor contraction of the lam grid with the unit components; the reference's
        # Cholesky test of S + P (pygam.py:499-537) runs as one batch, its retry loop only for the failures
        S = np.eye(m) * np.sqrt(EPS)
        gams, lams, grids = [], [], []
        for param_grid in param_grid_list:
            try:
                gam = self._clone_for_candidate()
                gam.set_params(**param_grid)
                gam._validate_params()
                gams.append(gam)
                lams.append(np.asarray(_flatten(gam.terms.lam), dtype=np.float64))
                grids.append(param_grid)
            except ValueError as error:
                if self.verbose:
                    warnings.warn(str(error) + "\non model with params:\n" + str(param_grid) + "\nskipping...\n")
        if not gams:
            return True
        L = np.stack(lams)
        scratch = copy.deepcopy(self.terms)
        unit_n, unit_r = [], []
        for c in range(L.shape[1]):
            e = np.zeros(L.shape[1])
            e[c] = 1.0
            scratch.lam = list(e)
            pn, pr = scratch._penalties(split=True)
            unit_n.append(pn)
            unit_r.append(pr)
        Pn_all = (L @ np.stack(unit_n).reshape(L.shape[1], m * m)).reshape(-1, m, m)
        Pr_all = (L @ np.stack(unit_r).reshape(L.shape[1], m * m)).reshape(-1, m, m)
        A_all = S[None] + (Pn_all + Pr_all)
        ok = eng.chol_ok_batch(A_all)
        cands = []
        for k, gam in enumerate(gams):
            try:
                EtE = A_all[k] if ok[k] else gam._cholesky_check(eng, A_all[k])
                cands.append((gam, EtE - Pn_all[k], Pn_all[k], grids[k]))
            except ValueError as error:
                if self.verbose:
                    warnings.warn(str(error) + "\non model with params:\n" + str(grids[k]) + "\nskipping...\n")
        K = len(cands)
        if K == 0:
            return True
        rank, nranks = world()
        lo, hi = K * rank // nranks, K * (rank + 1) // nranks
        loc = eng.solve_batch(np.stack([c[1] for c in cands[lo:hi]]) if hi > lo else np.zeros((0, m, m)),
                              np.stack([c[2] for c in cands[lo:hi]]) if hi > lo else np.zeros((0, m, m)),
                              want_cov=True) if hi > lo else dict(coef=np.zeros((0, m)), edof=np.zeros(0),
                                                                  info=np.zeros(0), cov=np.zeros((0, m, m)))
        packed = np.concatenate([loc["coef"], loc["edof"][:, None], np.asarray(loc["info"], dtype=np.float64)[:, None],
                                 loc["cov"].reshape(hi - lo, m * m)], axis=1)
        packed = np.where(np.isfinite(packed), packed, 0.0) if nranks > 1 else packed
        full = all_gather_rows(packed, K, lo, data.device) if nranks > 1 else packed
        prev = np.asarray(models[-1].coef_, dtype=np.float64) if models else np.ones(m) * np.sqrt(EPS)
        new_models, new_scores = [], []
        for k, (gam, _, _, param_grid) in enumerate(cands):
            beta, edof, info = full[k, :m], float(full[k, m]), int(full[k, m + 1])
            if info != 0 or not np.isfinite(beta).all():
                if self.verbose:
                    warnings.warn("penalised normal equations are not positive definite\non model with params:\n"
                                  + str(param_grid) + "\nskipping...\n")
                continue
            dev = dev_at(beta)
            if not (dev > 1e-7 * yy):
                return False  # cancellation: take the data passes instead
            gam._n_global = n
            gam.statistics_ = _LazyStats(gam, {"n_samples": n, "m_features": data.n_features},
                                         lazy=("edof_per_coef", "p_values"))
            gam.statistics_._ctx = (G, cands[k][1] + cands[k][2])
            if not hasattr(gam, "logs_"):
                gam.logs_ = {}
            for c in gam.callbacks:
                gam.logs_.setdefault(c, [])
            diff = float(np.linalg.norm(prev - beta) / np.linalg.norm(beta))
            gam.coef_ = prev
            gam._log_start(dev_at(prev), 0.0, n)
            gam._log_end(diff)
            gam.coef_ = beta.copy()
            if diff >= gam.tol and gam.max_iter >= 2:
                gam._log_start(dev, 0.0, n)
                diff = 0.0  # the second solve of the same system returns the same coefficients
                gam._log_end(diff)
            gam._linear_statistics_from_moments(n, edof, dev, yy, sy, full[k, m + 2:].reshape(m, m))
            if diff >= gam.tol:
                print("did not converge")
            new_models.append(gam)
            new_scores.append(gam.statistics_[objective])
            prev = gam.coef_
        models.extend(new_models)
        scores.extend(new_scores)
        for k in _DEVICE_KEYS:
            for gam in new_models:
                gam.__dict__[k] = self.__dict__.get(k)
        return True

    def _linear_statistics_from_moments(self, n, edof, dev, yy, sy, cov_unscaled):
        """_estimate_model_statistics (pygam.py:993-1057) of a Normal / identity model without weights from the
        residual sum of squares `dev` and the moments of y: the log-likelihood of Normal(mu, scale) rows
        (distributions.py:110-131) is -dev / (2 scale^2) - n (log(2 pi) / 2 + log(scale)), the null model's mean
        is sum(y) / n (pygam.py:1141-1147)."""
        dist, st = self.distribution, self.statistics_
        st["edof"] = edof
        if not dist._known_scale:
            dist.scale = float(np.sqrt(dev / (n - edof)))
        st["scale"] = dist.scale
        st["cov"] = cov_unscaled * dist.scale ** 2
        st["se"] = np.maximum(st["cov"].diagonal(), 0.0) ** 0.5
        s2 = dist.scale ** 2
        const = n * (0.5 * np.log(2.0 * np.pi) + np.log(dist.scale))
        ll = -0.5 * dev / s2 - const
        null_dev = yy - sy * sy / n
        null_ll = -0.5 * null_dev / s2 - const
        st["AIC"] = -2 * ll + 2 * edof + 2 * (not dist._known_scale)
        st["AICc"] = st["AIC"] + 2 * (edof + 1) * (edof + 2) / (n - edof - 2)
        st["pseudo_r2"] = OrderedDict([("explained_deviance", 1.0 - dev / null_dev), ("McFadden", ll / null_ll),
                                       ("McFadden_adj", 1.0 - (ll - edof) / null_ll)])
        st["GCV"], st["UBRE"] = self._estimate_GCV_UBRE(dev, n, edof)
        st["loglikelihood"] = ll
        st["deviance"] = dev / s2


class _ShapeOnly:
    """stand-in for X in Term.compile(): only the shape is needed once the column extrema and
    level counts come from the device"""

    def __init__(self, n, F):
        self.shape = (n, F)


def _flatten_knots(terms):
    out = []
    for t in terms:
        if t.istensor:
            out.extend(q.edge_knots_ for q in t._terms)
        elif not t.isintercept:
            out.append(getattr(t, "edge_knots_", None))
    return out


class LinearGAM(GAM):
    """Normal distribution, identity link (pygam.py:2371-2506)"""

    def __init__(self, terms="auto", max_iter=100, tol=1e-4, scale=None, callbacks=("deviance", "diffs"),
                 fit_intercept=True, verbose=False, **kwargs):
        self.scale = scale
        super().__init__(terms=terms, distribution="normal", link="identity", max_iter=max_iter, tol=tol,
                         callbacks=callbacks, fit_intercept=fit_intercept, verbose=verbose, **kwargs)

    def _validate_params(self, check_terms=True):
        if isinstance(self.distribution, str):
            self.distribution = Distribution("normal", scale=self.scale)
        super()._validate_params(check_terms=check_terms)

    def prediction_intervals(self, X, width=0.95, quantiles=None):
        """pygam.py:2477-2506"""
        if not self._is_fitted:
            raise AttributeError("GAM has not been fitted. Call fit first.")
        return self._get_quantiles(X, width, quantiles, prediction=True)


class LogisticGAM(GAM):
    """Binomial distribution, logit link (pygam.py:2509-2682)"""

    def __init__(self, terms="auto", max_iter=100, tol=1e-4, callbacks=("deviance", "diffs", "accuracy"),
                 fit_intercept=True, verbose=False, **kwargs):
        super().__init__(terms=terms, distribution="binomial", link="logit", max_iter=max_iter, tol=tol,
                         callbacks=callbacks, fit_intercept=fit_intercept, verbose=verbose, **kwargs)

    def accuracy(self, X=None, y=None, mu=None):
        """pygam.py:2609-2650"""
        if mu is None:
            mu = self.predict_mu(X)
        y = np.asarray(y).ravel()
        return ((mu > 0.5).astype(int) == y).sum() / len(y)

    def score(self, X, y):
        return self.accuracy(X=X, y=y)

    def predict(self, X):
        """pygam.py:2652-2666"""
        return self.predict_mu(X) > 0.5

    def predict_proba(self, X):
        return self.predict_mu(X)


class PoissonGAM(GAM):
    """Poisson distribution, log link, exposure support (pygam.py:2685-3046)"""

    def __init__(self, terms="auto", max_iter=100, tol=1e-4, callbacks=("deviance", "diffs"),
                 fit_intercept=True, verbose=False, **kwargs):
        super().__init__(terms=terms, distribution="poisson", link="log", max_iter=max_iter, tol=tol,
                         callbacks=callbacks, fit_intercept=fit_intercept, verbose=verbose, **kwargs)

    @staticmethod
    def _exposure_to_weights(y, exposure=None, weights=None):
        """pygam.py:2838-2890 incl. the float32 casts"""
        y = np.asarray(y, dtype=np.float64).ravel()
        if exposure is not None:
            exposure = np.array(exposure).astype("f").ravel()
            if not np.isfinite(exposure).all():
                raise ValueError("sample exposure data must not contain Inf nor NaN")
        else:
            exposure = np.ones_like(y).astype("float64")
        if len(exposure) != len(y):
            raise ValueError(f"Inconsistent input data shapes. found {len(y)} and {len(exposure)}")
        y = y / exposure
        if weights is not None:
            weights = np.array(weights).astype("f").ravel()
            if not np.isfinite(weights).all():
                raise ValueError("sample weights data must not contain Inf nor NaN")
        else:
            weights = np.ones_like(y).astype("float64")
        if len(exposure) != len(weights):
            raise ValueError(f"Inconsistent input data shapes. found {len(weights)} and {len(exposure)}")
        return y, weights * exposure

    def fit(self, X, y=None, exposure=None, weights=None):
        if isinstance(X, DeviceData) and exposure is None:
            return super().fit(X, y, weights)
        y, weights = self._exposure_to_weights(y, exposure, weights)
        return super().fit(X, y, weights)

    def predict(self, X, exposure=None):
        """pygam.py:2922-2959"""
        mu = self.predict_mu(X)
        if exposure is not None:
            exposure = np.array(exposure).astype("f")
        else:
            exposure = np.ones(len(mu)).astype("f")
        if len(exposure) != len(mu):
            raise ValueError(f"Inconsistent input data shapes. found {len(mu)} and {len(exposure)}")
        return mu * exposure

    def gridsearch(self, X, y=None, exposure=None, weights=None, return_scores=False, keep_best=True,
                   objective="auto", **param_grids):
        """pygam.py:2961-3046"""
        if not (isinstance(X, DeviceData) and exposure is None):
            y, weights = self._exposure_to_weights(y, exposure, weights)
        return super().gridsearch(X, y, weights=weights, return_scores=return_scores, keep_best=keep_best,
                                  objective=objective, **param_grids)


class GammaGAM(GAM):
    """Gamma distribution, log link (pygam.py:3049-3165)"""

    def __init__(self, terms="auto", max_iter=100, tol=1e-4, scale=None, callbacks=("deviance", "diffs"),
                 fit_intercept=True, verbose=False, **kwargs):
        self.scale = scale
        super().__init__(terms=terms, distribution="gamma", link="log", max_iter=max_iter, tol=tol,
                         callbacks=callbacks, fit_intercept=fit_intercept, verbose=verbose, **kwargs)


class InvGaussGAM(GAM):
    """Inverse Gaussian distribution, log link (pygam.py:3168-3284)"""

    def __init__(self, terms="auto", max_iter=100, tol=1e-4, scale=None, callbacks=("deviance", "diffs"),
                 fit_intercept=True, verbose=False, **kwargs):
        self.scale = scale
        super().__init__(terms=terms, distribution="inv_gauss", link="log", max_iter=max_iter, tol=tol,
                         callbacks=callbacks, fit_intercept=fit_intercept, verbose=verbose, **kwargs)


class ExpectileGAM(GAM):
    """Normal distribution, identity link, asymmetric least squares (pygam.py:3287-3551)"""

    def __init__(self, terms="auto", max_iter=100, tol=1e-4, scale=None, callbacks=("deviance", "diffs"),
                 fit_intercept=True, expectile=0.5, verbose=False, **kwargs):
        self.scale = scale
        self.expectile = expectile
        super().__init__(terms=terms, distribution="normal", link="identity", max_iter=max_iter, tol=tol,
                         callbacks=callbacks, fit_intercept=fit_intercept, verbose=verbose, **kwargs)
        self._expectile = float(expectile)

    def _validate_params(self, check_terms=True):
        if not (0 < self.expectile < 1):
            raise ValueError(f"expectile must be in (0,1), but found {self.expectile}")
        self._expectile = float(self.expectile)
        if isinstance(self.distribution, str):
            self.distribution = Distribution("normal", scale=self.scale)
        super()._validate_params(check_terms=check_terms)

    def _get_quantile_ratio(self, X, y):
        """fraction of the targets below the fitted curve (pygam.py:3462-3480)"""
        y_pred = self.predict(X)
        return (y_pred > np.asarray(y).ravel()).mean()

    def fit_quantile(self, X, y, quantile, max_iter=20, tol=0.01, weights=None):
        """bisection on `expectile` until the model quantile matches (pygam.py:3482-3551)"""
        if quantile <= 0 or quantile >= 1:
            raise ValueError(f"quantile must be on (0, 1), but found {quantile}")
        if tol <= 0:
            raise ValueError(f"tol must be float > 0 {tol}")
        if max_iter <= 0:
            raise ValueError(f"max_iter must be int > 0 {max_iter}")
        if not self._is_fitted:
            self.fit(X, y, weights=weights)
        max_ = 1.0
        min_ = 0.0
        n_iter = 0
        while n_iter < max_iter:
            ratio = self._get_quantile_ratio(X, y)
            if np.abs(ratio - quantile) < tol:
                break
            if ratio < quantile:
                min_ = self.expectile
            else:
                max_ = self.expectile
            expectile = (max_ + min_) / 2.0
            self.set_params(expectile=expectile)
            self.fit(X, y, weights=weights)
            n_iter += 1
        if n_iter == max_iter and self.verbose:
            warnings.warn("maximum iterations reached")
        return self
