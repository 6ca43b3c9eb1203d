#!/usr/bin/env python
"""Per-source-line instruction counts from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`
(first profiled kernel instance).  usage: ncu_lines.py file.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
agg = {}
first_func = None
n_func = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        if first_func is None:
            first_func = r[1]
    elif r[0] == "Line No":
        hdr = r
        iE = hdr.index("Instructions Executed")
        iS = hdr.index("# Samples")
    elif r[0] == "Kernel Name":
        n_func += 1
        if n_func > 1:
            break
    elif len(r) > 10 and r[2] == "-" and r[0].isdigit():
        try:
            e = int(r[iE]); s = int(r[iS])
        except ValueError:
            continue
        key = (cur_file, int(r[0]), r[1].strip()[:90])
        a = agg.setdefault(key, [0, 0])
        a[0] += e
        a[1] += s
tot = sum(v[0] for v in agg.values())
print("total warp-instructions:", tot)
byfile = {}
for (f, l, s), v in agg.items():
    byfile[f] = byfile.get(f, 0) + v[0]
for f, v in sorted(byfile.items(), key=lambda x: -x[1]):
    print("%6.2f%%  %s" % (100.0 * v / tot, f))
for (f, l, s), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print("%6.2f%% samp=%6d %s:%d  %s" % (100.0 * v[0] / tot, v[1], f, l, s))
