V=$PWD/build/variants/librvmc_b200_omufu.so
for lib in "" $V; do
  echo "== lib=${lib:-default}"
  RVMC_B200_LIB=$lib timeout 120 python tools/parity_margin.py
  RVMC_B200_LIB=$lib timeout 200 python bench.py --steps 30 --no-cpu-baseline --no-pipe-peaks > gpurun_out/r02l_ab.json 2> gpurun_out/r02l_ab.err
  python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02l_ab.json").read().strip().splitlines()[-1])
print("value %.4e ms %.4f e2e %.4e grid %.4f indep %.4f C3(0.7) %.4e C4 %.4e" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["grid"]["seconds"], d["grid_independent"]["seconds"], d["configs"]["C3"]["variants"][0]["binary_rvs_per_s"], d["configs"]["C4"]["binary_rvs_per_s"]))
PY
done
RVMC_B200_LIB=$V timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
