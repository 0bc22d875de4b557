#!/bin/bash
# round 2, call 5 (8 GPUs): NVLink exchange check on 2 GPUs, then BASELINE configs[3] at its stated size on 1/2/4/8 GPUs
mkdir -p gpurun_out
set -x
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_exchange_check.py > gpurun_out/r2c5_dist_check.log 2>&1
echo "rc=$?" >> gpurun_out/r2c5_dist_check.log
run() {  # N, tag, extra env
  n=$1; tag=$2; shift 2
  if [ "$n" = "1" ]; then
    env "$@" timeout 600 python bench.py --gpus 1 --scale 1.0 --steps 5 --warmup 2 --no-citeulike > gpurun_out/r2c5_scaled1_${tag}.json 2> gpurun_out/r2c5_scaled1_${tag}.err
  else
    env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 296$n$n bench.py --gpus $n --scale 1.0 --steps 5 --warmup 2 --no-citeulike > gpurun_out/r2c5_scaled1_${tag}.json 2> gpurun_out/r2c5_scaled1_${tag}.err
  fi
}
run 8 n8 X=1
run 4 n4 X=1
run 2 n2 X=1
run 1 n1 X=1
run 8 n8_nosplit HYPREC_SPLIT=0
