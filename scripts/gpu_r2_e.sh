#!/bin/bash
# round 2, fifth GPU session: C5 level-2 diagnostics, the whole bench line on one GPU, sampled slabs, C1 kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "sampled" > gpurun_out/r2e_tests_sampled.log 2>&1; tail -3 gpurun_out/r2e_tests_sampled.log
timeout 300 python scripts/probe_c5_two_level.py 1e-6 > gpurun_out/r2e_c5_1e-6.log 2>&1; tail -1 gpurun_out/r2e_c5_1e-6.log | cut -c1-900
timeout 300 python scripts/probe_c1.py > gpurun_out/r2e_c1.log 2>&1; tail -3 gpurun_out/r2e_c1.log
timeout 1500 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2e_bench.json; tail -5 gpurun_out/r2e_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2e_bench_ref.json 2>&1; tail -c 600 gpurun_out/r2e_bench_ref.json
