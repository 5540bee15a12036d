"""Host-side input validation helpers with the names and behaviour of the reference's ``pygam/utils.py``.

``check_array`` (utils.py:118-201), ``check_y`` (204-248), ``check_X`` (251-336), ``check_X_y`` (339-355),
``check_iterable_depth`` (866-896), ``make_2d`` (96-115), ``flatten`` (846-863), ``sig_code`` (569-590),
``round_to_n_decimal_places`` (470-495) and ``nice_repr`` (core.py:7-92).  They validate
HOST arrays a caller is about to hand to a model (type, shape, sample count, finiteness, link domain, categorical
range); the models run the same checks on the device once the rows are in HBM (``GAM._prepare_data``, ``_check_y``,
``_check_predict_X``).  No n-sized arithmetic of a fit goes through this module.
"""

import numpy as np

SKSPIMPORT = False  # the reference's switch for scikit-sparse (utils.py:8-14): the m x m solve here is pgb_solve


def make_2d(array, verbose=True):
    """utils.py:96-115: a 1-d array becomes a column"""
    array = np.asarray(array)
    if array.ndim < 2:
        msg = f"Expected 2D input data array, but found {array.ndim}D. Expanding to 2D."
        if verbose:
            import warnings

            warnings.warn(msg)
        array = np.atleast_1d(array)[:, None]
    return array


def check_array(array, force_2d=False, n_feats=None, ndim=None, min_samples=1, name="Input data", verbose=True):
    """utils.py:118-201: dimensionality, numeric type, sample count, feature count, finiteness"""
    if force_2d:
        array = make_2d(array, verbose=verbose)
        ndim = 2
    else:
        array = np.array(array)
    dtype = array.dtype
    if dtype.kind not in ["i", "f"]:
        try:
            array = array.astype("float")
        except ValueError as e:
            raise ValueError(f"{name} must be type int or float, but found type: {dtype}\n"
                             f"Try transforming data with a LabelEncoder first.") from e
    if ndim is not None and array.ndim != ndim:
        raise ValueError(f"{name} must have {ndim} dimensions. found shape {array.shape}")
    n = array.shape[0]
    if n < min_samples:
        raise ValueError(f"{name} should have at least {min_samples} samples, but found {n}")
    if n_feats is not None:
        m = array.shape[1]
        if m != n_feats:
            raise ValueError(f"{name} must have {n_feats} features, but found {m}")
    if not np.isfinite(array).all():
        raise ValueError(f"{name} must not contain Inf nor NaN")
    return array


def _link_domain(link, dist):
    name, levels = getattr(link, "name", str(link)), float(getattr(dist, "levels", 1) or 1)
    return {"logit": (0.0, levels), "log": (0.0, np.inf)}.get(name, (-np.inf, np.inf))


def check_y(y, link, dist, min_samples=1, verbose=True):
    """utils.py:204-248: a finite numeric vector inside the domain of the link"""
    y = np.ravel(y)
    y = check_array(y, force_2d=False, min_samples=min_samples, ndim=1, name="y data", verbose=verbose)
    lo, hi = _link_domain(link, dist)
    if np.any(y < lo) or np.any(y > hi):
        raise ValueError(f"y data is not in domain of {getattr(link, 'name', link)} link function. "
                         f"Expected domain: {[float(lo), float(hi)]}, but found {[float(y.min()), float(y.max())]}")
    return y


def check_X(X, n_feats=None, min_samples=1, edge_knots=None, dtypes=None, features=None, verbose=True):
    """utils.py:251-336: a finite numeric (n, F) array; categorical features inside the range seen at fit time"""
    if bool(features):
        features = flatten(features)
        max_feat = max(features)
        if n_feats is None:
            n_feats = max_feat
        n_feats = max(n_feats, max_feat)
    X = check_array(X, force_2d=True, n_feats=n_feats, min_samples=min_samples, name="X data", verbose=verbose)
    if edge_knots is not None and dtypes is not None and features is not None:
        for feat, ek, dt in zip(features, edge_knots, dtypes):
            if dt == "categorical":
                lo, hi = float(np.min(ek)), float(np.max(ek))
                col = X[:, feat]
                if col.min() < lo or col.max() > hi:
                    raise ValueError(f"X data is out of domain for categorical feature {feat}. "
                                     f"Expected data on [{lo}, {hi}], but found data on [{col.min()}, {col.max()}]")
    return X


def check_X_y(X, y):
    """utils.py:339-355"""
    if len(X) != len(y):
        raise ValueError(f"Inconsistent input and output data shapes. found X: {np.shape(X)} and y: {np.shape(y)}")


def check_iterable_depth(obj, max_depth=100):
    """utils.py:866-896: how deeply nested an iterable is (strings do not count)"""

    def _iterable(o):
        return hasattr(o, "__iter__") and not isinstance(o, str)

    depth = 0
    while depth < max_depth and _iterable(obj) and len(obj) > 0:
        obj = obj[0] if hasattr(obj, "__getitem__") else next(iter(obj))
        depth += 1
    return depth


def flatten(iterable):
    """utils.py:866-896: nested iterables become one flat list; anything else is returned as it is"""
    if not (hasattr(iterable, "__iter__") and not isinstance(iterable, str)):
        return iterable
    out = []
    for el in list(iterable):
        el = flatten(el)
        out.extend(el if isinstance(el, list) else [el])
    return out


def round_to_n_decimal_places(array, n=3):
    """utils.py:470-495: floats rounded to n decimals for printing; a float whose str() is already of the form 1e-05
    is left alone"""
    if isinstance(array, float) and "%.e" % array == str(array):
        return array
    shape = np.shape(array)
    return ((np.atleast_1d(array) * 10 ** n).round().astype("int") / (10.0 ** n)).reshape(shape)


def nice_repr(name, param_kvs, line_width=30, line_offset=5, decimals=3, args=None, flatten_attrs=True):
    """core.py:7-92: ``name(key=value, ...)`` with the keys in alphabetical order, floats rounded, and a new line
    (indented by line_offset) whenever the next ``key=value,`` would pass line_width"""
    if not param_kvs and not args:
        return f"{name}()"
    items = [(None, a) for a in (args or [])] + sorted(param_kvs.items(), key=lambda kv: kv[0])
    lines, line = [], name + "("
    for pos, (k, v) in enumerate(items):
        if flatten_attrs and k != "terms":
            v = flatten(v)
        text = str(round_to_n_decimal_places(v, n=decimals)) if isinstance(v, (float, np.ndarray)) else repr(v)
        piece = f"{text}," if k is None else f"{k}={text},"
        if len(line + piece) <= line_width:
            line += piece
        else:
            lines.append(line)
            line = " " * line_offset + piece
        if len(line) < line_width and pos + 1 < len(items):
            line += " "
    lines.append(line[:-1] + ")")  # without the trailing comma
    return "\n".join(lines)


def sig_code(p_value):
    """utils.py:569-590"""
    assert 0 <= p_value <= 1, "p_value must be on [0, 1]"
    for cut, code in ((0.001, "***"), (0.01, "**"), (0.05, "*"), (0.1, ".")):
        if p_value < cut:
            return code
    return " "
