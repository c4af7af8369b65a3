#!/bin/bash
# round 2, session P: level-1 thread-per-box evaluation kernel (parity), C5 at 1e-8 against the hand-over factor
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_two_level.py tests/test_gpu_parity.py -x -q -m gpu -k "two_level or single" > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2p_tests.log
timeout 300 python bench.py --workload c5 --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r2p_bench_c5.json 2> gpurun_out/r2p_bench_c5.err; echo "bench c5 rc=$?"; cut -c1-600 gpurun_out/r2p_bench_c5.json; tail -2 gpurun_out/r2p_bench_c5.err
for sf in 4096 256 16; do
C5_SWITCH_FACTOR=$sf timeout 300 python scripts/probe_c5_two_level.py 1e-8 > gpurun_out/r2p_c5_1e-8_sf$sf.log 2>&1; echo "c5 sf=$sf rc=$?"; cut -c1-900 gpurun_out/r2p_c5_1e-8_sf$sf.log
done
