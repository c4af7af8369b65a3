#!/bin/bash
# ncu evidence pass (run AFTER the plain commands exited 0): launch list of a bench step + captures of the hot kernels.
# Reports are exported to CSV on the box and deleted (gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1: $(( $(date +%s) - t0 )) s"; }
BENCH1="python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline"
$BENCH1 > gpurun_out/bench_1step.json 2> gpurun_out/bench_1step.err || exit 1
lap bench_plain
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4_full.csv $BENCH1 > gpurun_out/ncu_launches.log 2>&1
lap launch_list
export_rep() {   # $1 = report stem
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/$1.details.txt 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
}
FULL="ncu --set full --clock-control none --import-source on"
PK="python scripts/profile_kernel.py hcub genz_product_peak 262144"
$PK > gpurun_out/pk_prod.log 2>&1 && $FULL -k regex:hcub_pair -s 1 -c 1 -o gpurun_out/hcub_pair_prod_c4 -f $PK > gpurun_out/ncu_prod.log 2>&1
export_rep hcub_pair_prod_c4; lap c4_prod_full
PK="python scripts/profile_kernel.py quadgk 65536"
$PK > gpurun_out/pk_quadgk.log 2>&1 && $FULL -k regex:quadgk -s 1 -c 1 -o gpurun_out/quadgk_c1 -f $PK > gpurun_out/ncu_quadgk.log 2>&1
export_rep quadgk_c1; lap c1
PK="python scripts/profile_kernel.py fixed genz_oscillatory 16384"
$PK > gpurun_out/pk_fixed.log 2>&1 && $FULL -k regex:tensor_kernel -s 1 -c 1 -o gpurun_out/tensor_c3 -f $PK > gpurun_out/ncu_fixed.log 2>&1
export_rep tensor_c3; lap c3
PK="python scripts/profile_kernel.py sampled 1048576"
$PK > gpurun_out/pk_sampled.log 2>&1 && $FULL -k regex:strided_kernel -s 2 -c 2 -o gpurun_out/sampled_c2 -f $PK > gpurun_out/ncu_sampled.log 2>&1
export_rep sampled_c2; lap c2
PK="python scripts/probe_c5.py 22"
$PK > gpurun_out/pk_c5.log 2>&1 && $FULL -k regex:eval_kernel -s 100 -c 1 -o gpurun_out/single_eval_c5 -f $PK > gpurun_out/ncu_c5.log 2>&1
export_rep single_eval_c5; lap c5
du -sh gpurun_out; ls -la gpurun_out | head -40
