#!/bin/bash
# round 2, sixth GPU session: double-buffered select pass + closing sum, two-level status from the whole-integral test
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_two_level.py tests/test_gpu_rounds_batch.py -m gpu -q --tb=short > gpurun_out/r2f_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2f_tests.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/r2f_tests.log | tail -6
timeout 600 python scripts/probe_rb.py 262144 1 > gpurun_out/r2f_probe_full.log 2>&1; tail -8 gpurun_out/r2f_probe_full.log | cut -c1-260
timeout 300 python scripts/probe_c5_two_level.py 1e-7 > gpurun_out/r2f_c5_1e-7.log 2>&1; tail -1 gpurun_out/r2f_c5_1e-7.log | cut -c1-900
