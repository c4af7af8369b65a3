#!/bin/bash
# round 2, session S: BASELINE config 5 to its stated tolerance (reltol 1e-9, abstol 0) on ONE B200: two-level solve with retirement
mkdir -p gpurun_out
timeout 1500 python scripts/probe_c5_two_level.py 1e-9 > gpurun_out/r2s_c5_1e-9_n1.log 2>&1; echo "c5 1e-9 rc=$?"; cut -c1-1200 gpurun_out/r2s_c5_1e-9_n1.log
