# Final-build refresh of the 8-GPU row: bench at N = 8 and N = 1 on the same box.
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r02o_bench8.json 2> gpurun_out/r02o_bench8.err
timeout 250 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02o_bench1.json 2> gpurun_out/r02o_bench1.err
