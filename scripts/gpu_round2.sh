#!/bin/bash
# Round-2 opener (run under gpurun, one GPU): first measurements and ncu captures of the two paths built at the end of round 1
# (vector-valued HCubature, BatchIntegralFunction bridge).  Every command first runs plain (exit 0), then under ncu.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round2.sh'
set -u
mkdir -p gpurun_out
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
}
FULL="ncu --set full --clock-control none --import-source on"
LIST="ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv"
# 1. plain runs (numbers come from these, never from a run under ncu)
python scripts/profile_kernel.py hcubvec genz_gaussian 16384 3 4 > gpurun_out/r02_hcubvec_plain.log 2>&1 || exit 1
python scripts/profile_kernel.py bridge 4 18 > gpurun_out/r02_bridge_plain.log 2>&1 || exit 1
# 2. launch lists
$LIST --log-file gpurun_out/r02_launches_bridge.csv python scripts/profile_kernel.py bridge 4 18 > gpurun_out/r02_ncu_bridge_list.log 2>&1
# 3. full captures of the dominant kernels
$FULL -k regex:hcub_vec -s 1 -c 1 -o gpurun_out/r02_hcubvec_gauss -f python scripts/profile_kernel.py hcubvec genz_gaussian 16384 3 4 > gpurun_out/r02_ncu_hcubvec.log 2>&1
export_rep r02_hcubvec_gauss
$FULL -k regex:combine_kernel -s 20 -c 1 -o gpurun_out/r02_bridge_combine -f python scripts/profile_kernel.py bridge 4 18 > gpurun_out/r02_ncu_bridge_combine.log 2>&1
export_rep r02_bridge_combine
$FULL -k regex:points_kernel -s 20 -c 1 -o gpurun_out/r02_bridge_points -f python scripts/profile_kernel.py bridge 4 18 > gpurun_out/r02_ncu_bridge_points.log 2>&1
export_rep r02_bridge_points
ls -la gpurun_out/r02_*
