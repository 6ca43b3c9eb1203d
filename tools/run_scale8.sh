set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nproc; free -g | head -2
$TR --nproc-per-node 8 --master-port 29511 tools/d2h_bandwidth.py > gpurun_out/r02c_d2h_n8.json 2> gpurun_out/r02c_d2h_n8.err
$TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02c_bench8_bits.json 2> gpurun_out/r02c_bench8_bits.err
$TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --flag-transfer bytes --no-grid --no-configs > gpurun_out/r02c_bench8_bytes.json 2> gpurun_out/r02c_bench8_bytes.err
$TR --nproc-per-node 8 --master-port 29514 tools/multi_gpu_check.py > gpurun_out/r02c_mgc8.txt 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02c_bench1.json 2> gpurun_out/r02c_bench1.err
$TR --nproc-per-node 4 --master-port 29515 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02c_bench4.json 2> gpurun_out/r02c_bench4.err
$TR --nproc-per-node 2 --master-port 29516 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02c_bench2.json 2> gpurun_out/r02c_bench2.err
tail -3 gpurun_out/r02c_*.err
