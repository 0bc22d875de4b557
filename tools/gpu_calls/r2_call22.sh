#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2c22_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c22_gpu_tests.log
tail -4 gpurun_out/r2c22_gpu_tests.log
B="python bench.py --scale 0.25 --steps 3 --warmup 2 --no-citeulike --cpu-rows 4"
timeout 300 $B --eval-offset 120000 > gpurun_out/r2c22_new.json 2> gpurun_out/r2c22_new.err
HYPREC_SMALL_F64=0 timeout 300 $B --eval-offset 0 > gpurun_out/r2c22_nosmall.json 2> gpurun_out/r2c22_nosmall.err
python - <<'PY'
import json
for f in ("new", "nosmall"):
    try:
        d = json.loads(open("gpurun_out/r2c22_%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["rmse"], d["eval_ms"], d["eval_call_ms"], d["kernel_ms_per_step"], d.get("eval_topk", {}).get("fallback_users"))
    except Exception as e:
        print(f, "failed", e)
PY
CMD="python bench.py --scale 0.25 --steps 1 --warmup 1 --no-citeulike --cpu-rows 4 --eval-offset 120000"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2c22_launches_scaled025.csv $CMD > gpurun_out/r2c22_ncu.log 2>&1
tail -2 gpurun_out/r2c22_ncu.log
