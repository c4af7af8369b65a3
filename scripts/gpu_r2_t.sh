#!/bin/bash
# round 2, session T: the whole GPU suite (incl. parameter sensitivities, retirement, thread-per-box 6..8-D), C5 at 1e-8 with the automatic hand-over
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2t_tests.log | cut -c1-400
timeout 300 python scripts/probe_c5_two_level.py 1e-8 > gpurun_out/r2t_c5_1e-8.log 2>&1; echo "c5 rc=$?"; cut -c1-1100 gpurun_out/r2t_c5_1e-8.log
