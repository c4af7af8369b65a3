#!/bin/bash
# round 2, session W: two-level tests after the hand-over cap, then the default bench line (1 GPU) and the reference arm
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_two_level.py -q -m gpu > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2w_tests.log | cut -c1-300
timeout 1200 python bench.py > gpurun_out/r2w_bench_c4.json 2> gpurun_out/r2w_bench_c4.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2w_bench_c4.json; tail -2 gpurun_out/r2w_bench_c4.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2w_bench_ref.json 2>&1; tail -c 600 gpurun_out/r2w_bench_ref.json
