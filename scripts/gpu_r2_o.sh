#!/bin/bash
# round 2, session O: BASELINE config 5 to 1e-8 and to its stated tolerance 1e-9 on ONE GPU (two-level solve, thread-per-box 8-D rule)
mkdir -p gpurun_out
for rt in 1e-8 1e-9; do
timeout 900 python scripts/probe_c5_two_level.py $rt > gpurun_out/r2o_c5_$rt.log 2>&1; echo "c5 $rt rc=$?"; cut -c1-900 gpurun_out/r2o_c5_$rt.log
done
