#!/bin/bash
# per-kernel durations of a 512-point slice of the C5 grid (ncu launch list, serialised launches)
mkdir -p gpurun_out
python tools/prof_grid.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/grid_launches.csv python tools/prof_grid.py > gpurun_out/grid_ncu.log 2>&1
python - <<'PY'
import collections
import csv
rows = [r for r in csv.reader(l for l in open("gpurun_out/grid_launches.csv") if not l.startswith("=="))]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    if len(r) > iv:
        agg[r[ik].split("(")[0][-50:]].append(float(r[iv].replace(",", "")))
for k, v in agg.items():
    print("%-52s n=%3d  median %.1f us  total %.3f ms" % (k, len(v), sorted(v)[len(v) // 2] / 1e3, sum(v) / 1e6))
PY
