#!/bin/bash
# round 2, session N: rolled thread-per-box rule in 6..8-D: parity, C5 timings, ncu of the 8-D batch kernels (finite / two infinite axes)
mkdir -p gpurun_out
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
}
FULL="ncu --set full --clock-control none --import-source on"
timeout 900 python -m pytest tests/test_gpu_rounds_batch.py tests/test_gpu_two_level.py -x -q -m gpu -k "not c4_stated" > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2n_tests.log
for rt in 1e-6 1e-7; do
timeout 300 python scripts/probe_c5_two_level.py $rt > gpurun_out/r2n_c5_$rt.log 2>&1; echo "c5 rc=$?"; cut -c1-420 gpurun_out/r2n_c5_$rt.log
done
export PK_MAXEVALS=400000 PK_MODE=1
timeout 300 python scripts/profile_kernel.py hcub genz_gaussian 8192 8 > gpurun_out/r2n_pk.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/r2n_pk.log
timeout 900 $FULL -k regex:hcub_rb -s 1 -c 1 -o gpurun_out/r02_hcub_rb_gauss8d_batch -f python scripts/profile_kernel.py hcub genz_gaussian 8192 8 > gpurun_out/r2n_ncu2.log 2>&1; echo "ncu rc=$?"
export_rep r02_hcub_rb_gauss8d_batch
export PK_INF=1
timeout 300 python scripts/profile_kernel.py hcub genz_gaussian 8192 8 > gpurun_out/r2n_pk_inf.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/r2n_pk_inf.log
timeout 900 $FULL -k regex:hcub_rb -s 1 -c 1 -o gpurun_out/r02_hcub_rb_gauss8d_inf_batch -f python scripts/profile_kernel.py hcub genz_gaussian 8192 8 > gpurun_out/r2n_ncu3.log 2>&1; echo "ncu rc=$?"
export_rep r02_hcub_rb_gauss8d_inf_batch
