#!/bin/bash
mkdir -p gpurun_out
set -x
for mode in bulk fused; do
HYPREC_PUSH=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29688 bench.py --gpus 8 --scale 1.0 --steps 5 --warmup 2 --no-citeulike > gpurun_out/r2c17_scaled1_n8_$mode.json 2> gpurun_out/r2c17_scaled1_n8_$mode.err
done
