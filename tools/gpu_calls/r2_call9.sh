#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_bench_parity.py tests/test_gpu_parity.py tests/test_zz_device_extras.py -q -x > gpurun_out/r2c9_tests.log 2>&1
timeout 300 python bench.py --workload citeulike-a --steps 10 --warmup 3 > gpurun_out/r2c9_bench_cita.json 2> gpurun_out/r2c9_bench_cita.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c9_smoke_plain.log 2>&1 && \
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c9_memcheck_smoke.log 2>&1
echo "memcheck smoke rc=$?" >> gpurun_out/r2c9_memcheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -q -x -k "one_rank_exchange or tensor_memory_small_systems or eval_matches_oracle or device_initialised" > gpurun_out/r2c9_memcheck_tests.log 2>&1
echo "memcheck tests rc=$?" >> gpurun_out/r2c9_memcheck_tests.log
