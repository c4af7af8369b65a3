#!/bin/bash
# second GPU pass: full parity suite (incl. single-domain driver + golden fixtures), single-domain script at world 1, C5 probe
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python scripts/single_domain_mgpu.py > gpurun_out/single_w1.log 2>&1; echo "single rc=$?"; tail -5 gpurun_out/single_w1.log
timeout 600 python scripts/probe_c5.py 16 20 22 24 > gpurun_out/probe_c5.log 2>&1; echo "c5 rc=$?"; tail -8 gpurun_out/probe_c5.log
