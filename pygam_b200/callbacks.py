"""The reference's callback classes by name (``pygam/callbacks.py:98-236``).  The four built-in callbacks are computed
inside the passes of a fit (deviance and accuracy sums in the Gram pass, ``diff`` in the solve, the coefficients as they
are): ``callbacks=[Deviance(), Diffs()]`` is the same request as ``callbacks=['deviance', 'diffs']``.  A user-defined
``CallBack`` would receive the n-sized ``y`` / ``mu`` arrays of every iteration on the host (``pygam.py:780``), which this
engine never materialises there: it raises ``NotImplementedError`` at fit time."""


class CallBack:
    def __init__(self, name=None):
        self._name = name

    def __str__(self):
        return self._name if self._name is not None else self.__class__.__name__

    def __repr__(self):
        return f"{self.__class__.__name__}()"


class Deviance(CallBack):
    """deviance of the model at the start of every iteration (callbacks.py:111-142)"""

    def __init__(self):
        super().__init__(name="deviance")


class Accuracy(CallBack):
    """share of correctly predicted rows at the start of every iteration (callbacks.py:145-175)"""

    def __init__(self):
        super().__init__(name="accuracy")


class Diffs(CallBack):
    """relative change of the coefficients at the end of every iteration (callbacks.py:178-205)"""

    def __init__(self):
        super().__init__(name="diffs")


class Coef(CallBack):
    """the coefficients at the start of every iteration (callbacks.py:208-233)"""

    def __init__(self):
        super().__init__(name="coef")


CALLBACKS = {"deviance": Deviance, "diffs": Diffs, "accuracy": Accuracy, "coef": Coef}


def builtin_name(callback):
    """'deviance' for 'deviance' or Deviance(); None for anything that is not one of the four built-ins"""
    if isinstance(callback, str):
        return callback if callback in CALLBACKS else None
    for name, cls in CALLBACKS.items():
        if type(callback) is cls:
            return name
    return None
