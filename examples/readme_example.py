import sys; sys.path.insert(0,'.')
from rvmc_b200 import RVMC, BiasCorr            # or: sys.path.insert(0, "rvmc_b200/dropin"); from RVMC import RVMC
import numpy as np

cluster = RVMC(number_of_stars=10**5, seed=1)          # RVMC_example.py, unchanged apart from the optional seed
cluster.calculate_binary_velocities()
sigma = cluster.subsample(sample_size=12, measured_errors=1.2, number_of_samples=10**6)   # np.ndarray, km/s

t_obs = [np.array([0, 1, 7, 30, 180, 365.0]) * 86400] * 100                 # 100 stars, 6 epochs each
bc = BiasCorr(t_obs, number_of_samples=10**6, rv_errors=2.0, seed=2)        # 10^8 simulated binaries
bc.calculate_radial_velocities()
detected = bc.get_detected_binaries(delta_RV=20, n_sigma_RV=4)               # bool (100, 10**6), 3.5 ms end to end
p_detect = detected.mean(axis=1)

from rvmc_b200 import grid                                                   # findBestDistributions.py entry points
res = grid.grid_search(dict(period_pi=np.linspace(-0.95, 0.5, 64), mass_ratio_kappa=np.linspace(-0.95, 1, 64),
                            fbin=np.linspace(0.0625, 1, 16)), observed_sigma=25.0, number_of_samples=10**5)
print(res["best_params"])                                                    # 65 536 points in 0.13 s

print(sigma.shape, float(np.median(sigma)), detected.shape, float(p_detect.mean()))
