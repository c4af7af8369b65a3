#!/bin/bash
# round 2, 8-GPU session: the bench line with its multi-GPU secondaries, then BASELINE config 5 to its stated tolerance (two-level solve)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 1500 $TR bench.py --gpus $N --steps 1 --warmup 3 > gpurun_out/r2k_bench_n$N.json 2> gpurun_out/r2k_bench_n$N.err; echo "bench rc=$?"; tail -c 5000 gpurun_out/r2k_bench_n$N.json; tail -3 gpurun_out/r2k_bench_n$N.err
for rt in 1e-8 1e-9; do
  timeout 1200 $TR scripts/probe_c5_two_level.py $rt > gpurun_out/r2k_c5_${rt}_n$N.log 2>&1; grep '"rank": 0' gpurun_out/r2k_c5_${rt}_n$N.log | cut -c1-900
done
