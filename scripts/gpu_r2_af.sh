#!/bin/bash
# round 2, session AF: slow paths of the lock-step sin / sincos out of line (oscillatory Genz-Malik rule, QuadGK thread-per-problem kernel)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "quadgk or oscillatory or 2_pow_30" > gpurun_out/r2af_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2af_tests.log | cut -c1-400
timeout 300 python scripts/probe_variants.py 32768 2>&1 | tail -4
timeout 300 python scripts/probe_c1.py 2>&1 | tail -3
