mkdir -p gpurun_out
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
}
FULL="ncu --set full --clock-control none --import-source on"
export PK_MAXEVALS=100000
for fam in genz_oscillatory genz_product_peak genz_gaussian; do
  PK="python scripts/profile_kernel.py hcub $fam 262144"
  $FULL -k regex:hcub_pair -s 3 -c 1 -o gpurun_out/pair_${fam}_steady -f $PK > gpurun_out/ncu_$fam.log 2>&1
  export_rep pair_${fam}_steady
done
ls -la gpurun_out/*steady.raw.csv
