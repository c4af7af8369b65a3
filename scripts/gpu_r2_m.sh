#!/bin/bash
# round 2, session M: ncu captures of the 8-D thread-per-box round kernel (level 2 of C5, and a batch of 8-D Gaussians)
mkdir -p gpurun_out
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
}
FULL="ncu --set full --clock-control none --import-source on"
timeout 300 python scripts/probe_c5_two_level.py 1e-6 > gpurun_out/r2m_c5_1e-6.log 2>&1; echo "plain rc=$?"; cut -c1-700 gpurun_out/r2m_c5_1e-6.log
timeout 900 $FULL -k regex:hcub_rb -c 2 -o gpurun_out/r02_hcub_rb_c5_level2 -f python scripts/probe_c5_two_level.py 1e-6 > gpurun_out/r2m_ncu1.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2m_ncu1.log
export_rep r02_hcub_rb_c5_level2
export PK_MAXEVALS=400000 PK_MODE=1
timeout 300 python scripts/profile_kernel.py hcub genz_gaussian 8192 8 > gpurun_out/r2m_pk.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/r2m_pk.log
timeout 900 $FULL -k regex:hcub_rb -s 1 -c 1 -o gpurun_out/r02_hcub_rb_gauss8d_batch -f python scripts/profile_kernel.py hcub genz_gaussian 8192 8 > gpurun_out/r2m_ncu2.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2m_ncu2.log
export_rep r02_hcub_rb_gauss8d_batch
ls -la gpurun_out/
