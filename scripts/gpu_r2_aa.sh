#!/bin/bash
# round 2, session AA: first ncu captures of the vector-valued HCubature kernel, the sensitivity mode, and the bridge's two kernels
mkdir -p gpurun_out
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
}
FULL="ncu --set full --clock-control none --import-source on"
timeout 300 python scripts/profile_kernel.py hcubvec genz_gaussian 16384 3 4 > gpurun_out/r2aa_pk_vec.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/r2aa_pk_vec.log
timeout 900 $FULL -k regex:hcub_vec -s 1 -c 1 -o gpurun_out/r02_hcub_vec_gauss3d_nc4 -f python scripts/profile_kernel.py hcubvec genz_gaussian 16384 3 4 > gpurun_out/r2aa_ncu1.log 2>&1; echo "ncu rc=$?"
export_rep r02_hcub_vec_gauss3d_nc4
timeout 300 python scripts/profile_kernel.py bridge 4 18 > gpurun_out/r2aa_pk_bridge.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/r2aa_pk_bridge.log
timeout 900 $FULL -k regex:"points_kernel|combine_kernel" -s 60 -c 4 -o gpurun_out/r02_bridge_4d -f python scripts/profile_kernel.py bridge 4 18 > gpurun_out/r2aa_ncu2.log 2>&1; echo "ncu rc=$?"
export_rep r02_bridge_4d
ls -la gpurun_out/*.raw.csv | tail -3
