#!/bin/bash
# The ncu launch list of the default bench command (plain run first, as B200_PROFILING.md asks).
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
timeout 200 $CMD > gpurun_out/r02t_plain.json 2> gpurun_out/r02t_plain.err &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_bench_full.csv $CMD > gpurun_out/r02t_ncu.log 2>&1
wc -l gpurun_out/r02_launches_bench_full.csv
