"""Shim for `from astropy.constants import G` (RVMC.py:8, RVMC_bias_corr.py:6).

astropy >= 4.0 reports the CODATA-2018 value 6.6743e-11 m^3 kg^-1 s^-2; the product kernels pin
the same number (rvmc_b200/csrc/rvmc_constants.h)."""


class _Const:
    def __init__(self, value):
        self.value = value


G = _Const(6.6743e-11)
