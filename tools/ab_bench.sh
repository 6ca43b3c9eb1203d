#!/bin/bash
# A/B sweep of experimental library builds (tools/build_variant.py) on the headline kernel:
#   tools/ab_bench.sh base spt16 ...   -> gpurun_out/ab_<name>.json, one "name value ms" line each
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then lib=""; else lib="$PWD/build/variants/librvmc_b200_$v.so"; fi
  RVMC_B200_LIB=$lib python bench.py --no-grid --no-cpu-baseline --no-pipe-peaks --steps 40 --warmup 5 \
      > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/ab_%s.json" % v).read().strip().splitlines()[-1])
    print("%-14s value %.4e  ms %.4f  e2e %.4e  clk %s" % (v, d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"]))
except Exception as exc:
    print(v, "FAILED", exc)
PY
done
