"""`from RVMC_bias_corr import BiasCorr` for scripts written against the reference
(RVMC_bias_corr.py): put `rvmc_b200/dropin` on sys.path ahead of the reference checkout."""
from rvmc_b200.bias_corr import BiasCorr  # noqa: F401
