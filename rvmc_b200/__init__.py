"""rvmc_b200 -- B200-native Monte-Carlo hot path of FPABacks/RVMC behind the reference's own API.

`rvmc_b200.RVMC`, `rvmc_b200.BiasCorr` and `rvmc_b200.findBestDistributions` mirror
RVMC.py / RVMC_bias_corr.py / findBestDistributions.py; `rvmc_b200/dropin/` holds modules with
the reference's file names for `from RVMC import RVMC`-style scripts.
"""
__version__ = "0.1.0"

from ._abi import PopParams  # noqa: F401


def __getattr__(name):       # lazy: importing the package must not need torch / a GPU
    if name == "RVMC":
        from .rvmc import RVMC
        return RVMC
    if name == "plot_sigma_1d":
        from .rvmc import plot_sigma_1d
        return plot_sigma_1d
    if name == "BiasCorr":
        from .bias_corr import BiasCorr
        return BiasCorr
    if name == "findBestDistributions":
        from . import grid
        return grid
    raise AttributeError(name)
