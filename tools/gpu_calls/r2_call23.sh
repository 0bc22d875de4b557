#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_bench_parity.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2c23_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c23_gpu_tests.log
tail -4 gpurun_out/r2c23_gpu_tests.log
B="python bench.py --scale 0.25 --steps 3 --warmup 2 --no-citeulike --cpu-rows 4"
timeout 300 $B > gpurun_out/r2c23_new.json 2> gpurun_out/r2c23_new.err
python - <<'PY'
import json
for f in ("new",):
    try:
        d = json.loads(open("gpurun_out/r2c23_%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["rmse"], d["eval_ms"], d["eval_call_ms"], d["kernel_ms_per_step"], d.get("eval_topk", {}).get("fallback_users"))
    except Exception as e:
        print(f, "failed", e)
PY
