#!/bin/bash
# What is left of the GPU budget: the golden class runs and the late device extras on the rebuilt library.
O=gpurun_out/last4
mkdir -p $O
timeout 38 python -m pytest tests/test_gpu_parity.py tests/test_zz_device_extras.py -x -q -k "class_reproduces or one_byte or device_sweep" > $O/gpu_tests_subset.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests_subset.log
tail -3 $O/gpu_tests_subset.log
