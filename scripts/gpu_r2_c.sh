#!/bin/bash
# round 2, third GPU session: two-level single-domain solve (parity + C5 time-to-tolerance), rounds in 6-8 D, the whole GPU suite
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_two_level.py tests/test_gpu_rounds_batch.py -q --tb=short -s > gpurun_out/r2c_tests_new.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c_tests_new.log
tail -15 gpurun_out/r2c_tests_new.log
for rt in 1e-6 1e-7 1e-8; do
  timeout 600 python scripts/probe_c5_two_level.py $rt > gpurun_out/r2c_c5_$rt.log 2>&1; tail -2 gpurun_out/r2c_c5_$rt.log
done
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x --deselect tests/test_gpu_two_level.py --deselect tests/test_gpu_rounds_batch.py > gpurun_out/r2c_tests_all.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c_tests_all.log
tail -8 gpurun_out/r2c_tests_all.log
