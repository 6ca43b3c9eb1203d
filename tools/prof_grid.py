#!/usr/bin/env python
"""A 512-point slice of the C5 grid through the one-pass f_bin axis, for ncu launch lists."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rvmc_b200 import grid  # noqa: E402

torch.cuda.set_device(0)
axes = dict(period_pi=np.linspace(-0.95, 0.5, 16), mass_ratio_kappa=np.linspace(-0.95, 1.0, 16), fbin=np.linspace(0.0625, 1, 16))
for _ in range(2):
    r = grid.grid_search(axes, number_of_samples=100000, seed=1)
torch.cuda.synchronize()
print(r["shape"], float(r["stats"]["median"].mean()))
