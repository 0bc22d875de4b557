#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 300 python bench.py --scale 0.25 --steps 2 --warmup 1 --no-citeulike --cpu-rows 8 --eval-offset 0 > gpurun_out/r2c19_off0.json 2> gpurun_out/r2c19_off0.err
timeout 300 python bench.py --scale 0.25 --steps 2 --warmup 1 --no-citeulike --cpu-rows 8 --eval-offset 120000 > gpurun_out/r2c19_off120k.json 2> gpurun_out/r2c19_off120k.err
python - <<'PY'
import json
for f in ("off0", "off120k"):
    try:
        d = json.loads(open("gpurun_out/r2c19_%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["eval_ms"], d["eval_call_ms"], d["kernel_ms_per_step"], d.get("eval_topk"))
    except Exception as e:
        print(f, "failed", e)
PY
