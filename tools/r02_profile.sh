#!/bin/bash
# Round-2 ncu captures (one B200): plain run first, then the same command under ncu (B200_PROFILING.md).
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-grid --no-configs --no-cpu-baseline --no-pipe-peaks"
NCU="ncu --clock-control none"
timeout 120 $B > gpurun_out/r02p_plain_bench.log 2>&1 &&
timeout 300 $NCU --set full --import-source on -k regex:biascorr_fused -s 4 -c 1 -f -o gpurun_out/r02_biascorr_full $B > gpurun_out/r02p_ncu_bench.log 2>&1
timeout 120 python tools/prof_targets.py sigma > gpurun_out/r02p_plain_sigma.log 2>&1 &&
timeout 300 $NCU --set full --import-source on -k regex:sigma1d_fused -s 3 -c 1 -f -o gpurun_out/r02_sigma python tools/prof_targets.py sigma > gpurun_out/r02p_ncu_sigma.log 2>&1
timeout 120 python tools/prof_grid.py > gpurun_out/r02p_plain_grid.log 2>&1 &&
timeout 400 $NCU --set full --import-source on -k regex:"fbins|point_stats" -s 9 -c 3 -f -o gpurun_out/r02_grid python tools/prof_grid.py > gpurun_out/r02p_ncu_grid.log 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_launches_grid.csv python tools/prof_grid.py > gpurun_out/r02p_ncu_grid_list.log 2>&1
FULL="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-pipe-peaks --c3-orbits 1e8 --c4-realizations 1e5"
timeout 200 $FULL > gpurun_out/r02p_plain_full.log 2>&1 &&
timeout 500 $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv $FULL > gpurun_out/r02p_ncu_full.log 2>&1
ls -la gpurun_out/r02_*
