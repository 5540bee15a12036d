"""Term grammar of the GAM API: ``s()``, ``l()``, ``f()``, ``te()``, ``intercept`` and ``TermList``.

Mirrors the public surface of ``/root/reference/pygam/terms.py`` (constructors 497-599, 602-708,
711-956, 959-1101, 1104-1641, 1644-1979 and the shortcuts 1983-2046): same names, keyword
arguments, defaults and validation errors.  What differs is what a term *produces*: instead of
sparse model-matrix columns (``build_columns``) a compiled term flattens into one
``pgb_term`` record of the device model descriptor (``include/pgb.h``); the basis is then
evaluated on the fly inside the CUDA kernels.  Penalty and constraint matrices stay host-side
numpy, as in the reference.
"""

import copy
import warnings

import numpy as np
import scipy.linalg
import scipy.sparse

from . import penalties as _pen

DEFAULT_LAM = 0.6
MAX_ORDER = 7  # PGB_MAX_ORDER
MAX_MARGINALS = 4  # PGB_MAX_MARG


def _aslist(v):
    if isinstance(v, (list, tuple, np.ndarray)):
        return list(v)
    return [v]


def _term_get_params(obj, deep=False):
    """Core.get_params (core.py:139-167) for a term: the user-facing attributes (no leading or trailing underscore)
    minus the ones the reference hides per term kind (terms.py:535, 661, 845, 1026, 1304, 1722)"""
    attrs = dict(obj.__dict__)
    if deep:
        return attrs
    hidden = set(getattr(obj, "_hidden_params", ()))
    return {k: v for k, v in attrs.items() if k[0] != "_" and k[-1] != "_" and k not in hidden}


def _term_set_params(obj, deep=False, force=False, **parameters):
    """Core.set_params (core.py:169-192)"""
    names = _term_get_params(obj, deep=deep).keys()
    for k, v in parameters.items():
        if k in names or force or (hasattr(obj, k) and k == k.strip("_")):
            setattr(obj, k, v)
    return obj


def _term_build_columns(obj, X, verbose=False):
    """Term.build_columns / TermList.build_columns (terms.py:693-708, 910-957, 1077-1096, 1454-1477, 1901-1923): the
    columns of the model matrix of this (compiled) term for the rows of X, scipy sparse like the reference's.  The
    rows are evaluated on the device (pgb_model_matrix), the way every pass evaluates them."""
    from .engine import DeviceData, get_backend

    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    terms = obj if isinstance(obj, TermList) else TermList(obj)
    data = DeviceData.from_host(X)
    eng = get_backend().make_engine(terms, "normal", "identity", 1, None, data.n_features, data.device)
    return scipy.sparse.csc_matrix(eng.model_matrix(data).cpu().numpy())


class Term:
    """Common state of a term: feature, per-penalty ``lam``, ``penalties``, ``constraints``."""

    kind = "term"
    isintercept = False
    istensor = False

    def __init__(self, feature, lam=DEFAULT_LAM, penalties="auto", constraints=None, dtype="numerical",
                 by=None, verbose=False):
        self.feature = feature
        self.lam = lam
        self.penalties = penalties
        self.constraints = constraints
        self.dtype = dtype
        self.by = by
        self.verbose = verbose
        self.edge_knots_ = None
        self._validate()

    def __deepcopy__(self, memo):
        """terms hold scalars, strings, callables and (nested) lists of them, numpy arrays and -- te() -- marginal
        terms: copy exactly that, without the generic deep-copy machinery (grid search copies the term list once
        per candidate, pygam.py:2045)"""
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = _copy_value(v, memo)
        return new

    # -- argument normalisation (terms.py:148-210) --
    def _validate(self):
        if self.feature is not None:
            if not (isinstance(self.feature, (int, np.integer)) and self.feature >= 0):
                raise ValueError(f"feature must be an integer >= 0, but found {self.feature}")
            self.feature = int(self.feature)
        if self.by is not None:
            if not (isinstance(self.by, (int, np.integer)) and self.by >= 0):
                raise ValueError(f"by must be an integer >= 0, but found {self.by}")
            self.by = int(self.by)
        if self.dtype not in ("numerical", "categorical"):
            raise ValueError(f"dtype must be in ['numerical', 'categorical'], but found dtype = {self.dtype}")
        self.constraints = _aslist(self.constraints)
        for i, c in enumerate(self.constraints):
            if c is None:
                continue
            if not (callable(c) or c in _pen.CONSTRAINTS):
                raise ValueError(f"unknown constraint {c}")
        self.penalties = _aslist(self.penalties)
        for p in self.penalties:
            if not (callable(p) or p is None or p in _pen.PENALTIES):
                raise ValueError(f"penalties must be callable or in {list(_pen.PENALTIES)}, but found {p}")
        lam = _aslist(self.lam)
        if len(lam) == 1:
            lam = lam * len(self.penalties)
        if len(lam) != len(self.penalties):
            raise ValueError(f"expected 1 lam per penalty, but found lam = {self.lam}, "
                             f"penalties = {self.penalties}")
        for v in lam:
            if not (np.isfinite(v) and v >= 0):
                raise ValueError(f"lam must be >= 0 and finite, but found {v}")
        self.lam = [float(v) for v in lam]
        return self

    # -- grammar --
    def __add__(self, other):
        return TermList(self, other)

    def __radd__(self, other):
        return TermList(other, self)

    def __mul__(self, other):
        """terms.py:128-129"""
        raise NotImplementedError("terms cannot be multiplied: build an interaction with te()")

    def __eq__(self, other):
        """terms.py:117-120: terms are equal when their describing dicts are"""
        return isinstance(other, Term) and _info_equal(self.info, other.info)

    __hash__ = object.__hash__  # terms stay usable as dictionary keys (identity)

    @classmethod
    def build_from_info(cls, info):
        """terms.py:236-262, 1398-1420, 1790-1808: a term (or a term list) rebuilt from its ``info`` dict"""
        info = copy.deepcopy(info)
        cls_ = _term_classes()[info.pop("term_type")] if "term_type" in info else cls
        if cls_ is TermList:
            return TermList(*[Term.build_from_info(t) for t in info["terms"]])
        if cls_ is TensorTerm:
            return TensorTerm(*[Term.build_from_info(t) for t in info["terms"]], by=info.get("by"))
        import inspect
        accepted = set(inspect.signature(cls_.__init__).parameters) - {"self"}
        return cls_(**{k: v for k, v in info.items() if k in accepted})

    def __len__(self):
        return 1

    def __repr__(self):
        if self.isintercept:
            return "intercept"
        return f"{self.minimal_name}({self.feature})"

    minimal_name = "term"

    @property
    def info(self):
        d = {k: (list(v) if isinstance(v, (list, np.ndarray)) else v) for k, v in self.__dict__.items()
             if not k.endswith("_") and k != "verbose" and not k.startswith("_")}
        d["term_type"] = self.kind
        return d

    @property
    def n_coefs(self):
        raise NotImplementedError

    @property
    def hasconstraint(self):
        """terms.py:260-263 -- any entry that is not None counts, even the string "none" """
        return any(c is not None for c in self.constraints)

    def compile(self, X, verbose=False):
        return self

    _hidden_params = ("dtype", "constraints", "by", "edge_knots")

    def get_params(self, deep=False):
        return _term_get_params(self, deep=deep)

    def set_params(self, deep=False, force=False, **parameters):
        return _term_set_params(self, deep=deep, force=force, **parameters)

    def build_columns(self, X, verbose=False):
        return _term_build_columns(self, X, verbose=verbose)

    # -- m x m pieces --
    def _penalty_generators(self):
        """the (callable, lam) pairs of the term, 'auto' resolved (terms.py:322-340)"""
        out = []
        for pen, lam in zip(self.penalties, self.lam):
            if pen == "auto":
                if self.dtype == "numerical":
                    if self.kind == "spline_term":
                        pen = "periodic" if self.basis in ["cp"] else "derivative"
                    else:
                        pen = "l2"
                if self.dtype == "categorical":
                    pen = "l2"
            if pen is None:
                pen = "none"
            out.append((_pen.PENALTIES[pen] if isinstance(pen, str) else pen, lam))
        return out

    def build_penalties(self, verbose=False):
        """the penalty matrix of the term as the reference returns it (terms.py:307-349): scipy sparse, never densified
        (the reference's tests build it for a 10^6-coefficient te()).  The engine works on the dense form, ``_penalties``."""
        if self.isintercept:
            return scipy.sparse.csc_matrix((1, 1))
        n = self.n_coefs
        total = scipy.sparse.csc_matrix((n, n))
        for fn, lam in self._penalty_generators():
            total = total + scipy.sparse.csc_matrix(fn(n, None)) * lam
        return scipy.sparse.csc_matrix(total)

    def build_constraints(self, coef, constraint_lam, constraint_l2):
        """the constraint matrix of the term as the reference returns it (terms.py:351-396): scipy sparse"""
        if self.isintercept:
            return scipy.sparse.csc_matrix((1, 1))
        n = self.n_coefs
        total = scipy.sparse.csc_matrix((n, n))
        for c in self.constraints:
            if c is None:
                c = "none"
            fn = _pen.CONSTRAINTS[c] if isinstance(c, str) else c
            total = total + scipy.sparse.csc_matrix(fn(n, coef)) * constraint_lam
        if total.count_nonzero() > 0:
            total = total + constraint_l2 * scipy.sparse.identity(n, format="csc")
        return scipy.sparse.csc_matrix(total)

    def _penalties(self, split=False):
        """sum_k lam_k * penalty_k(n_coefs) (terms.py:307-349).  With split=True returns the sum
        as (null_part, rest): null_part collects the components that annihilate the constant
        vector of the block (difference penalties), see ``split_by_constant``."""
        if self.isintercept:
            return (np.array([[0.0]]), np.array([[0.0]])) if split else np.array([[0.0]])
        n = self.n_coefs
        total = np.zeros((n, n))
        parts = []
        for fn, lam in self._penalty_generators():
            comp = np.multiply(_pen.dense(fn, n, None), lam)
            total = total + comp
            parts.append(comp)
        return split_by_constant(parts, n) if split else total

    def _constraints(self, coef, constraint_lam, constraint_l2, split=False):
        """sum_k constraint_k(n, coef) * constraint_lam, plus constraint_l2 * I on a block
        with violations (terms.py:351-396)"""
        if self.isintercept:
            return (np.array([[0.0]]), np.array([[0.0]])) if split else np.array([[0.0]])
        n = self.n_coefs
        total = np.zeros((n, n))
        parts = []
        for c in self.constraints:
            if c is None:
                c = "none"
            fn = _pen.CONSTRAINTS[c] if isinstance(c, str) else c
            comp = _pen.dense(fn, n, coef) * constraint_lam
            total = total + comp
            parts.append(comp)
        if np.count_nonzero(total) > 0:
            total = total + constraint_l2 * np.eye(n)
            parts.append(constraint_l2 * np.eye(n))
        return split_by_constant(parts, n) if split else total


def _copy_value(v, memo):
    if isinstance(v, (list, tuple)):
        return type(v)(_copy_value(x, memo) for x in v)
    if isinstance(v, np.ndarray):
        return v.copy()
    if isinstance(v, Term):
        return copy.deepcopy(v, memo)
    if isinstance(v, dict):
        return {k: _copy_value(x, memo) for k, x in v.items()}
    return v  # int, float, str, bool, None, callables (penalty / constraint functions)


def split_by_constant(parts, n):
    """(sum of the components P with P @ 1 == 0, sum of the others).  Difference penalties and
    the D M D^T constraint penalties are integer matrices times a scalar, so P @ 1 is zero to
    rounding; an l2 component is not.  The solver zeroes the null block of the first sum after
    its change of basis (include/pgb.h, pgb_solve)."""
    null, rest = np.zeros((n, n)), np.zeros((n, n))
    ones = np.ones(n)
    for comp in parts:
        scale = np.abs(comp).max()
        if scale == 0.0:
            continue
        if np.abs(comp @ ones).max() <= 1e-13 * scale and np.abs(ones @ comp).max() <= 1e-13 * scale:
            null += comp
        else:
            rest += comp
    return null, rest


def _dense(a):
    return a.toarray() if hasattr(a, "toarray") else a


class Intercept(Term):
    kind = "intercept_term"
    isintercept = True
    minimal_name = "intercept"
    _hidden_params = ("by", "constraints", "dtype", "feature", "lam", "penalties", "edge_knots")

    def __init__(self, verbose=False):
        super().__init__(feature=None, lam=[0.0], penalties=[None], constraints=None, verbose=verbose)

    @property
    def n_coefs(self):
        return 1


class LinearTerm(Term):
    kind = "linear_term"
    minimal_name = "l"

    def __init__(self, feature, lam=DEFAULT_LAM, penalties="auto", verbose=False):
        super().__init__(feature=feature, lam=lam, penalties=penalties, constraints=None,
                         dtype="numerical", verbose=verbose)

    @property
    def n_coefs(self):
        return 1

    def compile(self, X, col_minmax=None, verbose=False):
        """terms.py:668-691: the column range is kept for generate_X_grid (pygam.py:1497-1500)"""
        if self.feature >= X.shape[1]:
            raise ValueError(f"term requires feature {self.feature}, but X has only {X.shape[1]} dimensions")
        if col_minmax is not None or hasattr(X, "__getitem__"):
            self.edge_knots_ = _edge_knots(X, self.feature, self.dtype, col_minmax)
        return self


class SplineTerm(Term):
    kind = "spline_term"
    minimal_name = "s"
    _bases = ["ps", "cp"]
    _hidden_params = ("edge_knots",)

    def __init__(self, feature, n_splines=20, spline_order=3, lam=DEFAULT_LAM, penalties="auto",
                 constraints=None, dtype="numerical", basis="ps", by=None, edge_knots=None,
                 verbose=False):
        self.basis = basis
        self.n_splines = n_splines
        self.spline_order = spline_order
        self.edge_knots = edge_knots
        super().__init__(feature=feature, lam=lam, penalties=penalties, constraints=constraints,
                         dtype=dtype, by=by, verbose=verbose)

    def _validate(self):
        super()._validate()
        if self.basis not in self._bases:
            raise ValueError(f"basis must be one of {self._bases}, but found: {self.basis}")
        self.n_splines = int(_check_int(self.n_splines, "n_splines", 0))
        self.spline_order = int(_check_int(self.spline_order, "spline_order", 0))
        if not self.n_splines > self.spline_order:
            raise ValueError(f"n_splines must be > spline_order. found: n_splines = {self.n_splines} "
                             f"and spline_order = {self.spline_order}")
        if self.spline_order > MAX_ORDER:
            raise NotImplementedError(f"spline_order > {MAX_ORDER} is not supported by the CUDA basis evaluation")
        if self.edge_knots is not None:
            self.edge_knots = np.asarray(self.edge_knots, dtype=np.float64).ravel()
        return self

    @property
    def n_coefs(self):
        return self.n_splines

    def compile(self, X, col_minmax=None, verbose=False):
        """terms.py:895-924: user edge knots are kept, otherwise [min, max] of the column"""
        if self.feature >= X.shape[1]:
            raise ValueError(f"term requires feature {self.feature}, but X has only {X.shape[1]} dimensions")
        if self.by is not None and self.by >= X.shape[1]:
            raise ValueError(f"by variable requires feature {self.by}, but X has only {X.shape[1]} dimensions")
        if self.edge_knots is not None:
            self.edge_knots_ = np.asarray(self.edge_knots, dtype=np.float64)
        if getattr(self, "edge_knots_", None) is None:
            self.edge_knots_ = _edge_knots(X, self.feature, self.dtype, col_minmax)
        # do the edge knots span the data of THIS fit?  edge_knots_ survives refits (terms.py:920) and can be set by
        # the user; an order-0 basis has an all-zero row outside them, so the block is a partition of unity -- and
        # (1 on the block, -1 on the intercept) a null vector of the design -- only if they do (null_vectors below)
        self._covers_data_ = None
        if col_minmax is not None or hasattr(X, "__getitem__"):
            if col_minmax is not None:
                lo, hi = col_minmax[self.feature]
            else:
                col = np.asarray(X[:, self.feature])
                lo, hi = col.min(), col.max()
            ek = np.sort(np.asarray(self.edge_knots_, dtype=np.float64))
            self._covers_data_ = bool(ek[0] <= lo and hi <= ek[-1])
        return self


class FactorTerm(SplineTerm):
    kind = "factor_term"
    minimal_name = "f"
    _encodings = ["one-hot", "dummy"]
    _hidden_params = ("basis", "by", "constraints", "dtype", "edge_knots", "n_splines", "spline_order")

    def __init__(self, feature, lam=DEFAULT_LAM, penalties="auto", coding="one-hot", verbose=False):
        self.coding = coding
        # n_splines is data dependent (terms.py:1071); 1 is a placeholder until compile()
        super().__init__(feature=feature, n_splines=1, spline_order=0, lam=lam, penalties=penalties,
                         constraints=None, dtype="categorical", basis="ps", by=None, verbose=verbose)

    def _validate(self):
        super()._validate()
        if self.coding not in self._encodings:
            raise ValueError(f"coding must be one of {self._encodings}, but found: {self.coding}")
        return self

    @property
    def n_coefs(self):
        return self.n_splines - (1 if self.coding == "dummy" else 0)

    def compile(self, X, col_minmax=None, n_levels=None, verbose=False):
        """terms.py:1054-1075: recomputed on every fit"""
        if self.feature >= X.shape[1]:
            raise ValueError(f"term requires feature {self.feature}, but X has only {X.shape[1]} dimensions")
        if n_levels is None:
            n_levels = len(np.unique(np.asarray(X[:, self.feature])))
        self.n_splines = int(n_levels)
        self.edge_knots_ = _edge_knots(X, self.feature, "categorical", col_minmax)
        return self


class TensorTerm(Term):
    kind = "tensor_term"
    minimal_name = "te"
    istensor = True
    _hidden_params = ("dtype", "feature", "edge_knots", "lam", "penalties", "constraints")
    _N_SPLINES = 10

    def __init__(self, *args, **kwargs):
        verbose = kwargs.pop("verbose", False)
        by = kwargs.pop("by", None)
        if "feature" in kwargs and args != tuple():
            raise ValueError("Marginal features in a TensorTerm must be specified either as `args` "
                             "or via keyword as `feature=...` but not both.")
        args = kwargs.pop("feature", args)
        m = len(args)
        if m < 2:
            raise ValueError("TensorTerm requires at least 2 marginal terms")
        if m > MAX_MARGINALS:
            raise NotImplementedError(f"tensor terms with more than {MAX_MARGINALS} marginals are not supported")
        for k, v in list(kwargs.items()):
            if isinstance(v, (list, tuple, np.ndarray)):
                if len(v) != m:
                    raise ValueError(f"Expected {k} to have length {m}, but found {k} = {v}")
            else:
                kwargs[k] = [v] * m
        margs = []
        for i, arg in enumerate(args):
            if isinstance(arg, TensorTerm):
                raise ValueError("TensorTerm does not accept other TensorTerms. "
                                 "Please build a flat TensorTerm instead of a nested one.")
            if isinstance(arg, Term):
                # s() and l() marginals (pygam/tests/test_terms.py:183-199: te(s(0), l(1)) is s(0, by=1) with one more
                # lam); f() marginals have no CUDA evaluation inside a tensor product
                if isinstance(arg, FactorTerm) or not isinstance(arg, (SplineTerm, LinearTerm)):
                    raise NotImplementedError("tensor marginals must be spline or linear terms")
                margs.append(arg)
                continue
            kw = {"n_splines": self._N_SPLINES}
            kw.update({k: v[i] for k, v in kwargs.items()})
            margs.append(SplineTerm(arg, **kw))
        self._terms = margs
        self.feature = [t.feature for t in margs]
        self.by = by
        self.verbose = verbose
        self.dtype = "numerical"
        if by is not None and not (isinstance(by, (int, np.integer)) and by >= 0):
            raise ValueError(f"by must be an integer >= 0, but found {by}")

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        new.__dict__.update({k: v for k, v in self.__dict__.items() if k != "_terms"})
        new._terms = [copy.deepcopy(t, memo) for t in self._terms]
        return new

    def __len__(self):
        return len(self._terms)

    def __getitem__(self, i):
        return self._terms[i]

    def __repr__(self):
        return "te(" + ", ".join(str(t.feature) for t in self._terms) + ")"

    @property
    def info(self):
        return {"term_type": self.kind, "by": self.by, "terms": [t.info for t in self._terms]}

    @property
    def n_coefs(self):
        return int(np.prod([t.n_coefs for t in self._terms]))

    @property
    def hasconstraint(self):
        return any(t.hasconstraint for t in self._terms)

    # the plural attributes of a tensor term are the lists over its marginals
    @property
    def lam(self):
        return [t.lam for t in self._terms]

    @property
    def edge_knots_(self):
        return [t.edge_knots_ for t in self._terms]

    def compile(self, X, col_minmax=None, verbose=False):
        for t in self._terms:
            t.compile(X, col_minmax=col_minmax, verbose=verbose)
        if self.by is not None and self.by >= X.shape[1]:
            raise ValueError(f"by variable requires feature {self.by}, but X has only {X.shape[1]} dimensions")
        return self

    def _marginal_penalty(self, i):
        """I x ... x P_i x ... x I, dense (terms.py:1499-1517)"""
        acc = None
        for j, t in enumerate(self._terms):
            Pj = t._penalties() if j == i else np.eye(t.n_coefs)
            acc = Pj if acc is None else np.kron(acc, Pj)
        return acc

    def _marginal_constraint(self, i, coef, constraint_lam, constraint_l2, split=False):
        """every 1-d slice of the coefficient tensor along marginal i gets that marginal's constraint matrix, dense
        (terms.py:1560-1609); split=True: (null part, rest)"""
        dims = [t.n_coefs for t in self._terms]
        size = int(np.prod(dims))
        C = np.zeros((size, size))
        Cn = np.zeros((size, size))
        idx = np.arange(size).reshape(dims)
        coef = np.asarray(coef)
        t = self._terms[i]
        slices = np.moveaxis(idx, i, 0).reshape(dims[i], size // dims[i])
        for sl in slices.T:
            if split:
                a, b = t._constraints(coef[sl], constraint_lam, constraint_l2, split=True)
                Cn[np.ix_(sl, sl)] += a
                C[np.ix_(sl, sl)] += b
            else:
                C[np.ix_(sl, sl)] += t._constraints(coef[sl], constraint_lam, constraint_l2)
        return (Cn, C) if split else C

    def _build_marginal_penalties(self, i):
        """I x ... x P_i x ... x I as a scipy sparse Kronecker product (terms.py:1499-1517)"""
        acc = None
        for j, t in enumerate(self._terms):
            Pj = t.build_penalties() if j == i else scipy.sparse.identity(t.n_coefs, format="csc")
            acc = Pj if acc is None else scipy.sparse.kron(acc, Pj, format="csc")
        return scipy.sparse.csc_matrix(acc)

    def _build_marginal_constraints(self, i, coef, constraint_lam, constraint_l2):
        """the constraint matrices of the 1-d slices along marginal i, assembled in coordinate form (terms.py:1560-1609)"""
        dims = [t.n_coefs for t in self._terms]
        size = int(np.prod(dims))
        coef = np.asarray(coef)
        idx = np.arange(size).reshape(dims)
        slices = np.moveaxis(idx, i, 0).reshape(dims[i], size // dims[i])
        data, rows, cols = [], [], []
        for sl in slices.T:
            Cs = self._terms[i].build_constraints(coef[sl], constraint_lam, constraint_l2).tocoo()
            data.append(Cs.data)
            rows.append(sl[Cs.row])
            cols.append(sl[Cs.col])
        return scipy.sparse.coo_matrix((np.concatenate(data), (np.concatenate(rows), np.concatenate(cols))),
                                       shape=(size, size)).tocsc()

    def build_penalties(self, verbose=False):
        """sum_i I x ... x P_i x ... x I, scipy sparse (terms.py:1479-1497)"""
        P = scipy.sparse.csc_matrix((self.n_coefs, self.n_coefs))
        for i in range(len(self._terms)):
            P = P + self._build_marginal_penalties(i)
        return scipy.sparse.csc_matrix(P)

    def build_constraints(self, coef, constraint_lam, constraint_l2):
        """the sum of the marginal constraint matrices, scipy sparse (terms.py:1519-1558)"""
        C = scipy.sparse.csc_matrix((self.n_coefs, self.n_coefs))
        for i in range(len(self._terms)):
            C = C + self._build_marginal_constraints(i, coef, constraint_lam, constraint_l2)
        return scipy.sparse.csc_matrix(C)

    def _penalties(self, split=False):
        """sum_i I x ... x P_i x ... x I (terms.py:1479-1517)"""
        parts = [self._marginal_penalty(i) for i in range(len(self._terms))]
        total = 0
        for acc in parts:
            total = total + acc
        return split_by_constant(parts, self.n_coefs) if split else total

    def _constraints(self, coef, constraint_lam, constraint_l2, split=False):
        """the sum of the marginal constraint matrices (terms.py:1519-1558)"""
        size = self.n_coefs
        C = np.zeros((size, size))
        Cn = np.zeros((size, size))
        for i in range(len(self._terms)):
            if split:
                a, b = self._marginal_constraint(i, coef, constraint_lam, constraint_l2, split=True)
                Cn += a
                C += b
            else:
                C += self._marginal_constraint(i, coef, constraint_lam, constraint_l2)
        return (Cn, C) if split else C


def _check_int(v, name, min_val):
    if isinstance(v, (bool,)) or not isinstance(v, (int, np.integer)) or v < min_val:
        if isinstance(v, (float, np.floating)) and float(v).is_integer() and v >= min_val:
            return int(v)
        raise ValueError(f"{name} must be int >= {min_val}, but found {name} = {v}")
    return v


def _edge_knots(X, feature, dtype, col_minmax=None):
    """gen_edge_knots (utils.py:592-621); col_minmax holds device-computed column extrema"""
    if col_minmax is not None:
        lo, hi = col_minmax[feature]
    else:
        col = np.asarray(X[:, feature])
        lo, hi = col.min(), col.max()
    if dtype == "categorical":
        return np.r_[lo - 0.5, hi + 0.5]
    knots = np.r_[lo, hi]
    if lo == hi:
        warnings.warn("Data contains constant feature. Consider removing and setting fit_intercept=True",
                      stacklevel=2)
    return knots


class TermList:
    """Ordered collection of unique terms (terms.py:1644-1979)."""

    _hidden_params = ()

    def get_params(self, deep=False):
        return _term_get_params(self, deep=deep)

    def set_params(self, deep=False, force=False, **parameters):
        return _term_set_params(self, deep=deep, force=force, **parameters)

    def build_columns(self, X, term=-1, verbose=False):
        """terms.py:1901-1923: the model matrix of all terms, or (term >= 0) the columns of one of them"""
        if term == -1:
            return _term_build_columns(self, X, verbose=verbose)
        return self._terms[term].build_columns(X, verbose=verbose)

    def __init__(self, *terms, **kwargs):
        self.verbose = kwargs.pop("verbose", False)
        if kwargs:
            raise ValueError(f"Unexpected keyword argument {kwargs.keys()}")
        seen, out = set(), []
        for t in terms:
            if isinstance(t, Term):
                items = [t]
            elif isinstance(t, TermList):
                items = t._terms
            else:
                raise ValueError(f"terms must be instances of Term or TermList, but found term: {t}")
            for it in items:
                key = repr(sorted(it.info.items(), key=lambda kv: kv[0]))
                if key in seen:
                    if self.verbose:
                        warnings.warn(f"skipping duplicate term: {it!r}")
                    continue
                seen.add(key)
                out.append(it)
        self._terms = out

    def __len__(self):
        return len(self._terms)

    def __getitem__(self, i):
        return self._terms[i]

    def __iter__(self):
        return iter(self._terms)

    def __add__(self, other):
        return TermList(self, other)

    def __radd__(self, other):
        return TermList(other, self)

    def __mul__(self, other):
        """terms.py:1756-1757"""
        raise NotImplementedError("terms cannot be multiplied: build an interaction with te()")

    def __eq__(self, other):
        """terms.py:1736-1739"""
        return isinstance(other, TermList) and _info_equal(self.info, other.info)

    __hash__ = object.__hash__

    @classmethod
    def build_from_info(cls, info):
        """terms.py:1790-1808"""
        return cls(*[Term.build_from_info(t) for t in copy.deepcopy(info)["terms"]])

    def __repr__(self):
        return " + ".join(repr(t) for t in self._terms)

    @property
    def info(self):
        return {"term_type": "term_list", "terms": [t.info for t in self._terms]}

    @property
    def n_coefs(self):
        return sum(t.n_coefs for t in self._terms)

    @property
    def hasconstraint(self):
        return any(t.hasconstraint for t in self._terms)

    def get_coef_indices(self, i=-1):
        """terms.py:1874-1899"""
        if i == -1:
            return list(range(self.n_coefs))
        if i >= len(self._terms):
            raise ValueError(f"requested {i}th term, but found only {len(self._terms)} terms")
        start = sum(t.n_coefs for t in self._terms[:i])
        return list(range(start, start + self._terms[i].n_coefs))

    def pop(self, i=None):
        if i is None:
            i = len(self) - 1
        if i >= len(self._terms) or i < 0:
            raise ValueError(f"requested pop {i}th term, but found only {len(self._terms)} terms")
        term = self._terms[i]
        self._terms = self._terms[:i] + self._terms[i + 1:]
        return term

    # ---- plural attributes (terms.py:399-494): values are consumed in term order ----
    def _slots(self):
        """the objects that own a ``lam`` list: non-intercept terms, tensor marginals flattened"""
        out = []
        for t in self._terms:
            if t.isintercept:
                continue
            if t.istensor:
                out.extend(t._terms)
            else:
                out.append(t)
        return out

    @property
    def lam(self):
        out = []
        for t in self._terms:
            if t.isintercept:
                continue
            out.append(t.lam)
        return out

    @lam.setter
    def lam(self, values):
        slots = self._slots()
        size = sum(len(s.lam) for s in slots)
        if np.ndim(values) == 0:
            flat = [values] * size
        else:
            flat = list(_flatten(values))
        if len(flat) != size:
            raise ValueError(f"Expected lam to have length {size}, but found lam = {values}")
        pos = 0
        for s in slots:
            k = len(s.lam)
            s.lam = flat[pos:pos + k]
            s._validate()
            pos += k

    def _set_plural(self, name, values):
        if name == "lam":
            self.lam = values
            return
        slots = [s for s in self._slots() if hasattr(s, name)]
        if np.ndim(values) == 0 or isinstance(values, str):
            flat = [values] * len(slots)
        else:
            flat = list(values)
        if len(flat) != len(slots):
            raise ValueError(f"Expected {name} to have length {len(slots)}, but found {name} = {values}")
        for s, v in zip(slots, flat):
            setattr(s, name, v)
            s._validate()

    @property
    def edge_knots_(self):
        return [t.edge_knots_ for t in self._terms if not t.isintercept and t.kind != "linear_term"]

    # ---- data-dependent state ----
    def compile(self, X, col_minmax=None, n_levels=None, verbose=False):
        for t in self._terms:
            if isinstance(t, FactorTerm):
                t.compile(X, col_minmax=col_minmax,
                          n_levels=None if n_levels is None else n_levels.get(t.feature), verbose=verbose)
            elif isinstance(t, (SplineTerm, TensorTerm, LinearTerm)):
                t.compile(X, col_minmax=col_minmax, verbose=verbose)
            else:
                t.compile(X, verbose=verbose)
        n_intercepts = sum(1 for t in self._terms if t.isintercept)
        if n_intercepts > 1:
            raise ValueError("cannot have more than one intercept")
        return self

    def features_used(self):
        feats = set()
        for t in self._terms:
            if t.isintercept:
                continue
            if t.istensor:
                feats.update(t.feature)
            else:
                feats.add(t.feature)
            if t.by is not None:
                feats.add(t.by)
        return feats

    # ---- m x m pieces ----
    def build_penalties(self, verbose=False):
        """block-diagonal penalty matrix as the reference returns it (terms.py:1925-1947): scipy sparse"""
        return scipy.sparse.block_diag([t.build_penalties() for t in self._terms], format="csc")

    def build_constraints(self, coef, constraint_lam, constraint_l2):
        """block-diagonal constraint matrix as the reference returns it (terms.py:1949-1979): scipy sparse"""
        coef = np.asarray(coef, dtype=np.float64).ravel()
        blocks, start = [], 0
        for t in self._terms:
            blocks.append(t.build_constraints(coef[start:start + t.n_coefs], constraint_lam, constraint_l2))
            start += t.n_coefs
        return scipy.sparse.block_diag(blocks, format="csc")

    def _penalties(self, split=False):
        """block-diagonal penalty matrix P (terms.py:1925-1947), dense (m, m); split=True returns
        (P_null, P_rest) with P = P_null + P_rest (see ``split_by_constant``)"""
        if split:
            parts = [t._penalties(split=True) for t in self._terms]
            return (scipy.linalg.block_diag(*[a for a, _ in parts]),
                    scipy.linalg.block_diag(*[b for _, b in parts]))
        return scipy.linalg.block_diag(*[t._penalties() for t in self._terms])

    def _constraints(self, coef, constraint_lam, constraint_l2, split=False):
        """block-diagonal constraint matrix C (terms.py:1949-1979), dense (m, m)"""
        coef = np.asarray(coef, dtype=np.float64).ravel()
        blocks, start = [], 0
        for t in self._terms:
            n = t.n_coefs
            blocks.append(t._constraints(coef[start:start + n], constraint_lam, constraint_l2, split=split))
            start += n
        if split:
            return (scipy.linalg.block_diag(*[a for a, _ in blocks]),
                    scipy.linalg.block_diag(*[b for _, b in blocks]))
        return scipy.linalg.block_diag(*blocks)

    def null_vectors(self):
        """Analytic null space of the design: every ps/cp spline, one-hot factor
        or tensor block without a ``by`` variable is a partition of unity, so
        (1 on the block, -1 on the intercept) -- or, without an intercept, (1 on the block, -1 on
        the first such block) -- is an exact null vector of X.  Returns (m, p)."""
        m = self.n_coefs
        blocks, icpt, start = [], None, 0
        for t in self._terms:
            n = t.n_coefs
            if t.isintercept:
                icpt = start
            elif t.by is None and (
                (isinstance(t, FactorTerm) and t.coding == "one-hot")
                or (isinstance(t, SplineTerm) and not isinstance(t, FactorTerm) and _rows_sum_to_one(t))
                or (t.istensor and all(_rows_sum_to_one(q) for q in t._terms))
            ):
                blocks.append((start, n))
            start += n
        vecs = []
        if icpt is not None:
            for s0, n in blocks:
                v = np.zeros(m)
                v[s0:s0 + n] = 1.0
                v[icpt] = -1.0
                vecs.append(v)
        elif len(blocks) > 1:
            b0, n0 = blocks[0]
            for s0, n in blocks[1:]:
                v = np.zeros(m)
                v[s0:s0 + n] = 1.0
                v[b0:b0 + n0] = -1.0
                vecs.append(v)
        if not vecs:
            return np.zeros((m, 0))
        return np.stack(vecs, axis=1)


def _rows_sum_to_one(t):
    """does every model-matrix row of the spline (marginal) sum to one on the data it was compiled for?  Orders
    >= 1 extrapolate linearly outside the edge knots (utils.py:744-763; the slopes of a partition of unity sum to
    zero), an order-0 basis is zero there: it qualifies only when the knots span the data (not known: only when
    the knots came from the data, i.e. the user gave none)."""
    if isinstance(t, LinearTerm):
        return False
    if t.spline_order > 0:
        return True
    covers = getattr(t, "_covers_data_", None)
    return t.edge_knots is None if covers is None else covers


def _term_classes():
    return {"term": Term, "intercept_term": Intercept, "linear_term": LinearTerm, "spline_term": SplineTerm,
            "factor_term": FactorTerm, "tensor_term": TensorTerm, "term_list": TermList}


def _info_equal(a, b):
    """dict equality that copes with the numpy arrays (user edge knots) an info dict may hold"""
    if isinstance(a, dict) and isinstance(b, dict):
        return a.keys() == b.keys() and all(_info_equal(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple, np.ndarray)) or isinstance(b, (list, tuple, np.ndarray)):
        if not (isinstance(a, (list, tuple, np.ndarray)) and isinstance(b, (list, tuple, np.ndarray))) or len(a) != len(b):
            return False
        return all(_info_equal(x, y) for x, y in zip(a, b))
    try:
        return bool(a == b)
    except Exception:  # noqa: BLE001
        return False


def _flatten(v):
    for x in v:
        if isinstance(x, (list, tuple, np.ndarray)):
            yield from _flatten(x)
        else:
            yield x


# shortcuts (terms.py:1983-2046)
def l(feature, lam=DEFAULT_LAM, penalties="auto", verbose=False):  # noqa: E743
    return LinearTerm(feature=feature, lam=lam, penalties=penalties, verbose=verbose)


def s(feature, n_splines=20, spline_order=3, lam=DEFAULT_LAM, penalties="auto", constraints=None,
      dtype="numerical", basis="ps", by=None, edge_knots=None, verbose=False):
    return SplineTerm(feature=feature, n_splines=n_splines, spline_order=spline_order, lam=lam,
                      penalties=penalties, constraints=constraints, dtype=dtype, basis=basis, by=by,
                      edge_knots=edge_knots, verbose=verbose)


def f(feature, lam=DEFAULT_LAM, penalties="auto", coding="one-hot", verbose=False):
    return FactorTerm(feature=feature, lam=lam, penalties=penalties, coding=coding, verbose=verbose)


def te(*args, **kwargs):
    return TensorTerm(*args, **kwargs)


intercept = Intercept()


def deepcopy_terms(terms):
    return copy.deepcopy(terms)
