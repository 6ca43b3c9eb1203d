"""`from RVMC import RVMC, plot_sigma_1d` for scripts written against the reference (RVMC.py):
put `rvmc_b200/dropin` on sys.path ahead of the reference checkout."""
from rvmc_b200.rvmc import RVMC, plot_sigma_1d  # noqa: F401
