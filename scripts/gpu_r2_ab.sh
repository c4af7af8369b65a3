#!/bin/bash
# round 2, session AB: C3 tensor rule with 4 rows per thread in the angle-addition loop: parity, time, ncu
mkdir -p gpurun_out
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_golden.py -q -m gpu -k "fixed or tensor or c3 or rule or golden or gauss_legendre or known" > gpurun_out/r2ab_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2ab_tests.log | cut -c1-300
timeout 300 python bench.py --workload c3 --no-secondary --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | cut -c1-330
timeout 300 python scripts/profile_kernel.py fixed genz_oscillatory 16384 > gpurun_out/r2ab_pk.log 2>&1; tail -1 gpurun_out/r2ab_pk.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tensor_kernel -s 1 -c 1 -o gpurun_out/r02_tensor_c3_rows4 -f python scripts/profile_kernel.py fixed genz_oscillatory 16384 > gpurun_out/r2ab_ncu.log 2>&1; echo "ncu rc=$?"
export_rep r02_tensor_c3_rows4
