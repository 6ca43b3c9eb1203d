set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 120 $TR --nproc-per-node 8 --master-port 29521 tools/grid_phases.py > gpurun_out/r02f_phases_n8.json 2> gpurun_out/r02f_phases_n8.err
timeout 120 $TR --nproc-per-node 8 --master-port 29522 tools/grid_phases.py --indep > gpurun_out/r02f_phases_indep_n8.json 2> gpurun_out/r02f_phases_indep_n8.err
timeout 200 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 20 --warmup 5 --no-configs > gpurun_out/r02f_bench8.json 2> gpurun_out/r02f_bench8.err
RVMC_HOST_THREADS=4 timeout 200 $TR --nproc-per-node 8 --master-port 29524 bench.py --gpus 8 --steps 20 --warmup 5 --no-configs --no-grid > gpurun_out/r02f_bench8_t4.json 2> gpurun_out/r02f_bench8_t4.err
RVMC_HOST_NO_AVX512=1 timeout 200 $TR --nproc-per-node 8 --master-port 29525 bench.py --gpus 8 --steps 20 --warmup 5 --no-configs --no-grid > gpurun_out/r02f_bench8_noavx.json 2> gpurun_out/r02f_bench8_noavx.err
timeout 100 $TR --nproc-per-node 8 --master-port 29526 tools/d2h_bandwidth.py > gpurun_out/r02f_d2h_n8.json 2> gpurun_out/r02f_d2h_n8.err
timeout 100 python bench.py --steps 20 --warmup 5 --no-configs --no-grid --no-cpu-baseline > gpurun_out/r02f_bench1.json 2> gpurun_out/r02f_bench1.err
