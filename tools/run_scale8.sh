# Round-2 scaling run on one 8 x B200 box: bench at N = 8, 4, 2, 1, grid phases, 8-GPU bit-identity check, reference arm.
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r02j_bench8.json 2> gpurun_out/r02j_bench8.err
timeout 100 $TR --nproc-per-node 8 --master-port 29542 tools/grid_phases.py > gpurun_out/r02j_phases_n8.json 2> gpurun_out/r02j_phases_n8.err
timeout 100 $TR --nproc-per-node 8 --master-port 29543 tools/multi_gpu_check.py > gpurun_out/r02j_mgc8.txt 2>&1
timeout 300 $TR --nproc-per-node 4 --master-port 29544 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/r02j_bench4.json 2> gpurun_out/r02j_bench4.err
timeout 300 $TR --nproc-per-node 2 --master-port 29545 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02j_bench2.json 2> gpurun_out/r02j_bench2.err
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/r02j_bench1.json 2> gpurun_out/r02j_bench1.err
timeout 100 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02j_ref.json 2> gpurun_out/r02j_ref.err
