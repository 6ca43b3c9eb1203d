set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 8 --master-port 29531 tools/d2h_bandwidth.py > gpurun_out/r02g_d2h_n8.json 2> gpurun_out/r02g_d2h_n8.err
timeout 120 $TR --nproc-per-node 8 --master-port 29532 tools/grid_phases.py > gpurun_out/r02g_phases_n8.json 2> gpurun_out/r02g_phases_n8.err
