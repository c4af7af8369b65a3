#!/bin/bash
# round 2, session Y: compiled user integrands (NVRTC) on the device; A/B of the C4 / C1 kernels against the build of commit 3ca76f0 ON THE SAME BOX
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_user_integrand.py -q -m gpu > gpurun_out/r2y_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2y_tests.log | cut -c1-400
FAMS=genz_gaussian,genz_c0,genz_product_peak,genz_oscillatory
echo "--- old (3ca76f0)"; (cd variants_old && timeout 300 python scripts/probe_rb.py 131072 1 $FAMS 2>&1 | grep "mode 1" | cut -c1-120; timeout 100 python scripts/probe_c1.py 2>&1 | tail -3)
echo "--- new";           (timeout 300 python scripts/probe_rb.py 131072 1 $FAMS 2>&1 | grep "mode 1" | cut -c1-120; timeout 100 python scripts/probe_c1.py 2>&1 | tail -3)
echo "--- old again";     (cd variants_old && timeout 300 python scripts/probe_rb.py 131072 1 $FAMS 2>&1 | grep "mode 1" | cut -c1-120)
