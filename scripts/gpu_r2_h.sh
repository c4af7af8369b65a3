#!/bin/bash
# round 2, GPU session H: thread-per-problem QuadGK, select early exit, fixed level-2 arenas
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_two_level.py tests/test_gpu_rounds_batch.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -k "quadgk or two_level or rounds or c1_ or golden" > gpurun_out/r2h_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2h_tests.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/r2h_tests.log | tail -8
timeout 300 python scripts/probe_c1.py > gpurun_out/r2h_c1.log 2>&1; tail -4 gpurun_out/r2h_c1.log
timeout 600 python scripts/probe_rb.py 262144 1 > gpurun_out/r2h_probe_full.log 2>&1; tail -8 gpurun_out/r2h_probe_full.log | cut -c1-200
timeout 300 python scripts/probe_c5_two_level.py 1e-7 > gpurun_out/r2h_c5_1e-7.log 2>&1; tail -1 gpurun_out/r2h_c5_1e-7.log | cut -c1-900
