"""``Core`` and ``nice_repr`` by the reference's names (``pygam/core.py``): the parameter / repr conventions shared by the
models and the terms.  Host-side bookkeeping only."""

from .utils import nice_repr  # noqa: F401  (re-exported: pygam.core.nice_repr)


class Core:
    """core.py:95-192: an object with a name, user-facing parameters (attributes without a leading or trailing
    underscore) and a wrapped repr"""

    def __init__(self, name=None, line_width=70, line_offset=3):
        self._name = name
        self._line_width = line_width
        self._line_offset = line_offset
        self._include = getattr(self, "_include", [])
        self._exclude = getattr(self, "_exclude", [])

    def __str__(self):
        return self._name if self._name is not None else self.__repr__()

    def __repr__(self):
        return nice_repr(self.__class__.__name__, self.get_params(), line_width=self._line_width,
                         line_offset=self._line_offset, decimals=4, args=None)

    def get_params(self, deep=False):
        attrs = dict(self.__dict__)
        for k in self._include:
            attrs[k] = getattr(self, k)
        if deep:
            return attrs
        return {k: v for k, v in attrs.items() if k[0] != "_" and k[-1] != "_" and k not in self._exclude}

    def set_params(self, deep=False, force=False, **parameters):
        names = self.get_params(deep=deep).keys()
        for k, v in parameters.items():
            if k in names or force or (hasattr(self, k) and k == k.strip("_")):
                setattr(self, k, v)
        return self
