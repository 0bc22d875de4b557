#!/bin/bash
# round 2, first GPU call: state of the tree as round 1 left it + the A/B measurements its captures point at
mkdir -p gpurun_out
set -x
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1
for cap in 0 159 127 95 63; do
  HYPREC_LOWRANK_CAP=$cap timeout 60 python tools/profile_step.py --epochs 4 --no-eval > gpurun_out/r2_cap_$cap.log 2>&1
done
timeout 400 python bench.py --workload scaled --scale 0.1 --eval-users 1024 --steps 2 --warmup 1 \
  > gpurun_out/r2_scaled01_eval.json 2> gpurun_out/r2_scaled01_eval.err
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
python - > gpurun_out/host.txt 2>&1 <<'PY'
import os, psutil
print(os.cpu_count(), psutil.virtual_memory().total/2**30)
PY
