#!/bin/bash
# round 2, first GPU session: parity of the round-based batched driver, mode 0 vs mode 1 throughput at the bench size,
# occupancy variants, one ncu --set full capture of the new kernel per family (product peak, oscillatory, Gaussian)
mkdir -p gpurun_out
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
}
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r2a_clocks.csv &
SMI=$!
timeout 900 python -m pytest tests/test_gpu_rounds_batch.py -q --tb=short -s > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
timeout 600 python scripts/probe_rb.py 262144 0,1 > gpurun_out/r2a_probe_full.log 2>&1
for w in 8 12 16; do
  timeout 300 python scripts/probe_rb.py 65536 1 genz_product_peak,genz_oscillatory,genz_gaussian $w > gpurun_out/r2a_probe_w$w.log 2>&1
done
kill $SMI
export PK_MODE=1
FULL="ncu --set full --clock-control none --import-source on"
for fam in genz_product_peak genz_oscillatory genz_gaussian; do
  PK="python scripts/profile_kernel.py hcub $fam 32768"
  timeout 300 $PK > gpurun_out/r2a_plain_$fam.log 2>&1 &&
  timeout 900 $FULL -k regex:hcub_rb -s 1 -c 1 -o gpurun_out/r2a_rb_$fam -f $PK > gpurun_out/r2a_ncu_$fam.log 2>&1
  export_rep r2a_rb_$fam
done
ls -la gpurun_out | tail -30
tail -5 gpurun_out/r2a_tests.log
cat gpurun_out/r2a_probe_full.log | tail -30
