#!/bin/bash
# A/B of experimental builds on the grid kernels: parity tests of the grid / fused kernels, then tools/ab_grid.py.
#   tools/ab_grid.sh base v1 v2 ...      (base = build/variants/librvmc_b200_base.so; results in gpurun_out/ab_grid_*.json)
mkdir -p gpurun_out
for v in "$@"; do
  lib="$PWD/build/variants/librvmc_b200_$v.so"
  if [ "$v" != base ]; then
    RVMC_B200_LIB=$lib timeout 120 python -m pytest tests/test_gpu_grid_and_dropin.py tests/test_gpu_pool_and_fused.py -m gpu -x -q > gpurun_out/ab_grid_$v.pytest.log 2>&1
    echo "$v pytest rc=$? $(tail -1 gpurun_out/ab_grid_$v.pytest.log)"
  fi
  RVMC_B200_LIB=$lib timeout 120 python tools/ab_grid.py $v > gpurun_out/ab_grid_$v.json 2> gpurun_out/ab_grid_$v.err
  echo "$v rc=$? $(cat gpurun_out/ab_grid_$v.json)"
done
