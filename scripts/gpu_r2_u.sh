#!/bin/bash
# round 2, session U: sensitivities + two-level tests after the automatic hand-over; C5 at 1e-8 with it
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_sensitivities.py tests/test_gpu_two_level.py tests/test_gpu_zz_batch_bridge.py tests/test_gpu_vector_cubature.py -q -m gpu > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2u_tests.log | cut -c1-600
timeout 300 python scripts/probe_c5_two_level.py 1e-8 > gpurun_out/r2u_c5_1e-8.log 2>&1; echo "c5 rc=$?"; cut -c1-1100 gpurun_out/r2u_c5_1e-8.log
