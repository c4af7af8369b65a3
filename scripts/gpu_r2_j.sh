#!/bin/bash
# GPU session J: the whole GPU suite with the current library, C1 kernels, C4 at the bench size
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2j_tests_all.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2j_tests_all.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/r2j_tests_all.log | tail -8
timeout 300 python scripts/probe_c1.py > gpurun_out/r2j_c1.log 2>&1; tail -4 gpurun_out/r2j_c1.log
timeout 600 python scripts/probe_rb.py 262144 1 > gpurun_out/r2j_probe_full.log 2>&1; tail -8 gpurun_out/r2j_probe_full.log | cut -c1-200
