# e2e at 8 GPUs only (short): bench without the grid / config legs.
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 150 $TR --nproc-per-node 8 --master-port 29571 bench.py --gpus 8 --steps 30 --warmup 5 --no-grid --no-configs > gpurun_out/r02r_bench8.json 2> gpurun_out/r02r_bench8.err
