"""The reference's link classes by name (``pygam/links.py:21-323``): ``GAM(link=LogLink(), ...)``.  The objects carry the
link's name -- which selects the CUDA implementation (``pgb_device.cuh``) -- and the reference's host-side methods
``link``, ``mu`` and ``gradient`` for arrays a caller handles himself."""

from .gam import Link


class IdentityLink(Link):
    def __init__(self):
        super().__init__("identity")


class LogitLink(Link):
    def __init__(self):
        super().__init__("logit")


class LogLink(Link):
    def __init__(self):
        super().__init__("log")


class InverseLink(Link):
    def __init__(self):
        super().__init__("inverse")


class InvSquaredLink(Link):
    def __init__(self):
        super().__init__("inv_squared")


LINKS = {"identity": IdentityLink, "logit": LogitLink, "log": LogLink, "inverse": InverseLink, "inv_squared": InvSquaredLink}
