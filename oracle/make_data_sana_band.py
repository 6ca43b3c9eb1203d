"""Medians and 16/84 percentiles of the six sigma_1D distributions in /root/reference/data_sana/*.dat
(100 000 values each, km/s) -> tests/golden/data_sana_band.json.  Build container only (reads
/root/reference).  No code in the reference reads these files and its current samplers do not reproduce
them (an older external simulation, medians 7-30 % apart); bench.py prints them beside
the C3 variants as a coarse sanity band, nothing more.

    python oracle/make_data_sana_band.py
"""
import json
import os

import numpy as np

SRC = "/root/reference/data_sana"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "data_sana_band.json")

band = {}
for name in sorted(os.listdir(SRC)):
    if not name.endswith(".dat"):
        continue
    x = np.loadtxt(os.path.join(SRC, name))
    band[name[:-4]] = {"n": int(x.size), "median": float(np.median(x)), "p16": float(np.percentile(x, 16)),
                       "p84": float(np.percentile(x, 84)), "mean": float(x.mean())}
with open(OUT, "w") as f:
    json.dump(band, f, indent=1, sort_keys=True)
print(json.dumps(band, indent=1))
