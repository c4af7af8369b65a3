#!/bin/bash
# first full GPU pass: parity tests, smoke, default bench, ncu launch list + full capture of the top kernel
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
SHORT="python bench.py --steps 1 --warmup 1 --nprob 8192 --no-secondary --no-cpu-baseline"
$SHORT > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
PK="python scripts/profile_kernel.py hcub genz_gaussian 4096"
$PK > gpurun_out/pk_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hcub_kernel -s 1 -c 1 -o gpurun_out/hcub_gauss_full -f $PK > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
