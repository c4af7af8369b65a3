#!/bin/bash
# round 2, session AH: compact (32-byte, dyadic) box records in the round-based kernel: on / off / hand-back forced at level 6
mkdir -p gpurun_out
timeout 300 python scripts/probe_variants.py 32768 > gpurun_out/r2ah_a.log 2>&1; tail -5 gpurun_out/r2ah_a.log
PV_COMPACT=0 timeout 300 python scripts/probe_variants.py 32768 > gpurun_out/r2ah_b.log 2>&1; tail -5 gpurun_out/r2ah_b.log
PV_LEVEL=6 timeout 300 python scripts/probe_variants.py 32768 > gpurun_out/r2ah_c.log 2>&1; tail -5 gpurun_out/r2ah_c.log
