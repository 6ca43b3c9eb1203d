#!/bin/bash
# Round-2 A/B sweep on one B200: launch geometry of the multi-epoch kernel, two library variants, the
# scratch size of the grid (rows per statistics launch) -> gpurun_out/r02d_*.json + a summary on stdout.
mkdir -p gpurun_out
run() {  # name, env..., -- bench args
  name=$1; shift
  envs=(); while [ $# -gt 0 ] && [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 180 python bench.py --no-cpu-baseline --no-pipe-peaks --no-configs --steps 40 --warmup 5 "$@" \
      > gpurun_out/r02d_$name.json 2> gpurun_out/r02d_$name.err
  python - "$name" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/r02d_%s.json" % v).read().strip().splitlines()[-1])
    g = d.get("grid") or {}
    gi = d.get("grid_independent") or {}
    print("%-16s value %.4e  ms %.4f  e2e %.4e  e2e+drv %.3e  grid %s  indep %s  clk %s" % (
        v, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["with_delta_rv"]["value"], g.get("seconds"),
        gi.get("seconds"), d["clocks"]["sm_mhz"]))
except Exception as exc:
    print(v, "FAILED", exc)
PY
}
V=$PWD/build/variants
run base -- --no-grid
run bytes -- --no-grid --flag-transfer bytes
run grid1184 --
run grid512 RVMC_GRID_SCRATCH_SLOTS=512 --
run grid256 RVMC_GRID_SCRATCH_SLOTS=256 --
run grid2368 RVMC_GRID_SCRATCH_SLOTS=2368 --
timeout 120 python tools/grid_phases.py > gpurun_out/r02d_phases_n1.json 2> gpurun_out/r02d_phases_n1.err
