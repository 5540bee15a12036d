"""The reference's distribution classes by name (``pygam/distributions.py:97-621``): ``GAM(distribution=GammaDist(), ...)``.
The objects carry the family's name, scale and levels -- which select the CUDA implementation of the variance, deviance
and log-likelihood (``pgb_device.cuh``) -- and the reference's host-side methods for arrays a caller handles himself."""

from .gam import Distribution


class NormalDist(Distribution):
    def __init__(self, scale=None):
        super().__init__("normal", scale=scale)


class BinomialDist(Distribution):
    def __init__(self, levels=1):
        super().__init__("binomial", levels=1 if levels is None else levels)


class PoissonDist(Distribution):
    def __init__(self):
        super().__init__("poisson")


class GammaDist(Distribution):
    def __init__(self, scale=None):
        super().__init__("gamma", scale=scale)


class InvGaussDist(Distribution):
    def __init__(self, scale=None):
        super().__init__("inv_gauss", scale=scale)


DISTRIBUTIONS = {"normal": NormalDist, "poisson": PoissonDist, "binomial": BinomialDist, "gamma": GammaDist,
                 "inv_gauss": InvGaussDist}
