#!/bin/bash
# round 2, session X: retirement out of line -- two-level / rounds tests, C4 step per family, C1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_two_level.py tests/test_gpu_rounds_batch.py -q -m gpu -k "not c4_stated" > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2x_tests.log | cut -c1-300
timeout 600 python bench.py --no-secondary --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/r2x_bench_c4.json 2> gpurun_out/r2x_bench_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2x_bench_c4.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], {k: round(v["kernel_ms"],1) for k,v in d["per_family"].items()})
PY
timeout 300 python bench.py --workload c1 --no-secondary --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | cut -c1-260
