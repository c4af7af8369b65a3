#!/bin/bash
# round 2, session AD: per-instruction stall samples (ncu source page) of the round-based kernel on the two heaviest C4 families
mkdir -p gpurun_out
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
}
FULL="ncu --set full --clock-control none --import-source on"
export PK_MODE=1
for fam in genz_product_peak genz_oscillatory; do
  PK="python scripts/profile_kernel.py hcub $fam 32768"
  timeout 200 $PK 2>&1 | tail -2
  timeout 400 $FULL -k regex:hcub_rb -s 1 -c 1 -o gpurun_out/r02b_rb_${fam} -f $PK > gpurun_out/r2ad_ncu_$fam.log 2>&1; echo "ncu rc=$?"
  export_rep r02b_rb_${fam}
done
ls -la gpurun_out/r02b_rb_*
