set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="bench.py --gpus 8 --steps 20 --warmup 5 --no-configs --no-grid"
timeout 250 $TR --nproc-per-node 8 --master-port 29531 tools/d2h_bandwidth.py > gpurun_out/r02g_d2h_n8.json 2> gpurun_out/r02g_d2h_n8.err
RVMC_HOST_THREADS=4 timeout 150 $TR --nproc-per-node 8 --master-port 29533 $B --flag-transfer bits > gpurun_out/r02g_b8_bits_t4.json 2> gpurun_out/r02g_b8_bits_t4.err
RVMC_HOST_THREADS=3 timeout 150 $TR --nproc-per-node 8 --master-port 29534 $B --flag-transfer bits > gpurun_out/r02g_b8_bits_t3.json 2> gpurun_out/r02g_b8_bits_t3.err
RVMC_HOST_THREADS=3 RVMC_FLAG_BYTE_SHARE=0.3 timeout 150 $TR --nproc-per-node 8 --master-port 29535 $B --flag-transfer hybrid > gpurun_out/r02g_b8_hyb30_t3.json 2> gpurun_out/r02g_b8_hyb30_t3.err
RVMC_HOST_THREADS=3 RVMC_FLAG_BYTE_SHARE=0.45 timeout 150 $TR --nproc-per-node 8 --master-port 29536 $B --flag-transfer hybrid > gpurun_out/r02g_b8_hyb45_t3.json 2> gpurun_out/r02g_b8_hyb45_t3.err
RVMC_HOST_THREADS=2 RVMC_FLAG_BYTE_SHARE=0.4 timeout 150 $TR --nproc-per-node 8 --master-port 29537 $B --flag-transfer hybrid > gpurun_out/r02g_b8_hyb40_t2.json 2> gpurun_out/r02g_b8_hyb40_t2.err
timeout 120 $TR --nproc-per-node 8 --master-port 29532 tools/grid_phases.py > gpurun_out/r02g_phases_n8.json 2> gpurun_out/r02g_phases_n8.err
timeout 300 $TR --nproc-per-node 8 --master-port 29538 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02g_bench8_full.json 2> gpurun_out/r02g_bench8_full.err
