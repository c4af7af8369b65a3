#!/bin/bash
# round 2, session R: C5 at 1e-8 with retirement, against the hand-over factor and the arena size
mkdir -p gpurun_out
C5_SWITCH_FACTOR=256 timeout 300 python scripts/probe_c5_two_level.py 1e-8 > gpurun_out/r2r_c5_1e-8_sf256.log 2>&1; echo "sf256 rc=$?"; cut -c1-1100 gpurun_out/r2r_c5_1e-8_sf256.log
C5_SUB_BOXES=262144 timeout 300 python scripts/probe_c5_two_level.py 1e-8 > gpurun_out/r2r_c5_1e-8_sub262144.log 2>&1; echo "sub262144 rc=$?"; cut -c1-1100 gpurun_out/r2r_c5_1e-8_sub262144.log
C5_SWITCH_FACTOR=1024 timeout 300 python scripts/probe_c5_two_level.py 1e-8 > gpurun_out/r2r_c5_1e-8_sf1024.log 2>&1; echo "sf1024 rc=$?"; cut -c1-1100 gpurun_out/r2r_c5_1e-8_sf1024.log
