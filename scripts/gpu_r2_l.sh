#!/bin/bash
# round 2, session L: thread-per-box Genz-Malik rule in 6..8 dimensions (gm_rule_thread3.cuh): parity, then C5 to 1e-7 on one GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rounds_batch.py tests/test_gpu_two_level.py -x -q -m gpu -k "not c4_stated" > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2l_tests.log
timeout 300 python scripts/probe_c5_two_level.py 1e-7 > gpurun_out/r2l_c5_1e-7.log 2>&1; echo "c5 rc=$?"; cut -c1-1200 gpurun_out/r2l_c5_1e-7.log
