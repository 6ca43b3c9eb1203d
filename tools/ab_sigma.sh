#!/bin/bash
# A/B sweep of experimental builds on the fused sigma_1D kernels (tools/bench_extra.py): C3 and the C5 slice
for v in "$@"; do
  if [ "$v" = base ]; then lib=""; else lib="$PWD/build/variants/librvmc_b200_$v.so"; fi
  RVMC_B200_LIB=$lib python tools/bench_extra.py 2>&1 | python -c "
import json, sys
out = {}
for line in sys.stdin:
    try:
        d = json.loads(line)
    except Exception:
        continue
    c = d['case']
    if c.startswith('C3 fbin0.70 Pmin1.4'): out['C3 f.70'] = d['binary_rvs_per_s']
    if c.startswith('C3 fbin0.12'): out['C3 f.12 slots'] = d['star_slots_per_s']
    if c.startswith('C4 N=332'): out['C4 N332'] = d['binary_rvs_per_s']
    if c.startswith('C5 grid'): out['C5 pts/s'] = d['points_per_s']
print('%-10s' % '$v', '  '.join('%s %.4e' % kv for kv in out.items()))
"
done
