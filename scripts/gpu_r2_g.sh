#!/bin/bash
# round 2, multi-GPU session (N ranks = $1): the single-domain path through NCCL (tests, two-level to tolerance, bit-identity with one GPU),
# the bench line with its multi-GPU secondaries, the reference arm under torchrun
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "two_gpus" > gpurun_out/r2g_tests_n$N.log 2>&1; tail -3 gpurun_out/r2g_tests_n$N.log
timeout 600 python scripts/probe_c5_two_level.py 1e-7 > gpurun_out/r2g_c5_1e-7_n1.log 2>&1; tail -1 gpurun_out/r2g_c5_1e-7_n1.log | cut -c1-500
timeout 600 $TR scripts/probe_c5_two_level.py 1e-7 > gpurun_out/r2g_c5_1e-7_n$N.log 2>&1; grep '"rank": 0' gpurun_out/r2g_c5_1e-7_n$N.log | cut -c1-500
timeout 1500 $TR bench.py --gpus $N --steps 1 --warmup 3 > gpurun_out/r2g_bench_n$N.json 2> gpurun_out/r2g_bench_n$N.err; echo "bench rc=$?"; tail -c 4500 gpurun_out/r2g_bench_n$N.json; tail -5 gpurun_out/r2g_bench_n$N.err
timeout 600 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2g_bench_ref_n$N.json 2>&1; tail -c 700 gpurun_out/r2g_bench_ref_n$N.json
