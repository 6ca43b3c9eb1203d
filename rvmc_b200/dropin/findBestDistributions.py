"""The grid-fit entry points of the reference's findBestDistributions.py (which is Python 2 and
cannot run) under their original names, GPU-backed: put `rvmc_b200/dropin` on sys.path."""
from rvmc_b200.grid import (create_sig1D_distribution, get_Pdistr_stats, get_fbinDist_stats, gaussian,  # noqa: F401
                            find_nearest, func, find_best_Pmin, find_best_fbin,
                            find_Pmin_fbin_individual_clulsters, run_bigGrid, grid_search, quad_func,
                            plot_Pmin_vs_age)
