#!/bin/bash
# round 2, multi-GPU session V (N ranks = $1): the bench line with its multi-GPU secondaries (N = 2: the c5 stated-config entry rehearsed at 1e-7)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
if [ "$N" -lt 8 ]; then export IB200_BENCH_C5_STATED_MIN_WORLD=$N IB200_BENCH_C5_STATED_RELTOL=1e-7; fi
timeout 1500 $TR bench.py --gpus $N --steps 1 --warmup 3 > gpurun_out/r2v_bench_n$N.json 2> gpurun_out/r2v_bench_n$N.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/r2v_bench_n$N.json; tail -3 gpurun_out/r2v_bench_n$N.err
