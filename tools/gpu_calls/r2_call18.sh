#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29644 bench.py --gpus 4 --scale 1.0 --steps 5 --warmup 2 --no-citeulike > gpurun_out/r2c18_scaled1_n4.json 2> gpurun_out/r2c18_scaled1_n4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --scale 1.0 --steps 5 --warmup 2 --no-citeulike > gpurun_out/r2c18_scaled1_n2.json 2> gpurun_out/r2c18_scaled1_n2.err
