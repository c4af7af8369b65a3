#!/bin/bash
# round 2, second GPU session: the cp.async-pipelined round kernel -- parity, throughput at the bench size, occupancy variants
# (launch bounds 2 / 3 / 4 blocks per SM), ncu --set full per family
mkdir -p gpurun_out
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
}
timeout 900 python -m pytest tests/test_gpu_rounds_batch.py tests/test_gpu_parity.py -q --tb=short -s -k "rounds or c4_stated or single" > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2b_tests.log
timeout 600 python scripts/probe_rb.py 262144 1 > gpurun_out/r2b_probe_full.log 2>&1
for v in minb4; do
  IB200_LIB=build_exp/lib_$v.so timeout 300 python scripts/probe_rb.py 65536 1 genz_product_peak,genz_oscillatory,genz_gaussian,genz_c0 > gpurun_out/r2b_probe_$v.log 2>&1
done
timeout 300 python scripts/probe_rb.py 65536 1 genz_product_peak,genz_oscillatory,genz_gaussian,genz_c0 > gpurun_out/r2b_probe_minb3.log 2>&1
export PK_MODE=1
FULL="ncu --set full --clock-control none --import-source on"
for fam in genz_product_peak genz_oscillatory genz_gaussian genz_c0 genz_corner_peak genz_discontinuous; do
  PK="python scripts/profile_kernel.py hcub $fam 32768"
  timeout 300 $PK > gpurun_out/r2b_plain_$fam.log 2>&1 &&
  timeout 900 $FULL -k regex:hcub_rb -s 1 -c 1 -o gpurun_out/r2b_rb_$fam -f $PK > gpurun_out/r2b_ncu_$fam.log 2>&1
  export_rep r2b_rb_$fam
done
tail -8 gpurun_out/r2b_tests.log
tail -12 gpurun_out/r2b_probe_full.log
