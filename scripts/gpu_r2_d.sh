#!/bin/bash
# round 2, fourth GPU session: thread-per-box rule v2 (shared sub-combinations, Newton reciprocal, constant-bank operands) -- parity,
# C4 throughput, ncu; C5 two-level time-to-tolerance on one GPU
mkdir -p gpurun_out
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
}
timeout 1200 python -m pytest tests/test_gpu_two_level.py tests/test_gpu_rounds_batch.py tests/test_golden.py -m gpu -q --tb=short -s > gpurun_out/r2d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2d_tests.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2d_tests.log | tail -8
timeout 600 python scripts/probe_rb.py 262144 1 > gpurun_out/r2d_probe_full.log 2>&1; tail -8 gpurun_out/r2d_probe_full.log
for rt in 1e-6 1e-7 1e-8; do
  timeout 900 python scripts/probe_c5_two_level.py $rt > gpurun_out/r2d_c5_$rt.log 2>&1; tail -1 gpurun_out/r2d_c5_$rt.log | cut -c1-600
done
export PK_MODE=1
FULL="ncu --set full --clock-control none --import-source on"
for fam in genz_product_peak genz_gaussian; do
  PK="python scripts/profile_kernel.py hcub $fam 32768"
  timeout 300 $PK > gpurun_out/r2d_plain_$fam.log 2>&1 &&
  timeout 900 $FULL -k regex:hcub_rb -s 1 -c 1 -o gpurun_out/r2d_rb_$fam -f $PK > gpurun_out/r2d_ncu_$fam.log 2>&1
  export_rep r2d_rb_$fam
done
