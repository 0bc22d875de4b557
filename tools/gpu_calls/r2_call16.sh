#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29688 bench.py --gpus 8 --scale 1.0 --steps 5 --warmup 2 --no-citeulike > gpurun_out/r2c16_scaled1_n8.json 2> gpurun_out/r2c16_scaled1_n8.err
