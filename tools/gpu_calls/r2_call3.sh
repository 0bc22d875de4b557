#!/bin/bash
# round 2, call 3 (2 GPUs): sharded tests on one GPU, two processes over CUDA IPC, bench at N = 1 and 2
mkdir -p gpurun_out
set -x
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_bench_parity.py -q -s > gpurun_out/r2c3_new_tests.log 2>&1
timeout 600 python bench.py --scale 0.1 --steps 3 --warmup 2 --no-citeulike > gpurun_out/r2c3_bench_scaled01_n1.json 2> gpurun_out/r2c3_bench_scaled01_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --scale 0.1 --steps 3 --warmup 2 --no-citeulike > gpurun_out/r2c3_bench_scaled01_n2.json 2> gpurun_out/r2c3_bench_scaled01_n2.err
