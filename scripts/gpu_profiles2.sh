#!/bin/bash
# second ncu pass, after the kernel changes of this round: launch list of one bench step + full captures of the changed kernels.
# The pair-mode captures use equal-cost problems (every problem capped at 1e5 evaluations): small box pools, so ncu's
# per-pass save/restore stays cheap (the bench-size capture of round 1 took 17 min for that reason).
mkdir -p gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1: $(( $(date +%s) - t0 )) s"; }
BENCH1="python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline"
$BENCH1 > gpurun_out/bench_1step_b.json 2> gpurun_out/bench_1step_b.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4_b.csv $BENCH1 > gpurun_out/ncu_launches_b.log 2>&1
lap launch_list
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
}
FULL="ncu --set full --clock-control none --import-source on"
export PK_MAXEVALS=100000
for fam in genz_oscillatory genz_product_peak genz_gaussian; do
  PK="python scripts/profile_kernel.py hcub $fam 262144"
  $PK > gpurun_out/pk_$fam.log 2>&1 && $FULL -k regex:hcub_pair -s 5 -c 1 -o gpurun_out/pair_${fam}_steady -f $PK > gpurun_out/ncu_$fam.log 2>&1
  export_rep pair_${fam}_steady; lap $fam
done
unset PK_MAXEVALS
PK="python scripts/profile_kernel.py fixed genz_oscillatory 16384"
$PK > gpurun_out/pk_fixed.log 2>&1 && $FULL -k regex:tensor_kernel -s 1 -c 1 -o gpurun_out/tensor_c3_b -f $PK > gpurun_out/ncu_fixed.log 2>&1
export_rep tensor_c3_b; lap c3
PK="python scripts/probe_c5.py 22"
$PK > gpurun_out/pk_c5.log 2>&1 && $FULL -k regex:eval_kernel -s 100 -c 1 -o gpurun_out/single_eval_c5_b -f $PK > gpurun_out/ncu_c5.log 2>&1
export_rep single_eval_c5_b; lap c5
du -sh gpurun_out
