#!/bin/bash
# round 2, session AR: the whole GPU suite, the default bench line and the reference arm on the final build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2ar_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ar_tests.log | cut -c1-300
( time timeout 1200 python bench.py > gpurun_out/r2ar_bench_c4.json 2> gpurun_out/r2ar_bench_c4.err ) 2> gpurun_out/r2ar_bench_time.txt; echo "bench rc=$?"; tail -2 gpurun_out/r2ar_bench_c4.err; grep real gpurun_out/r2ar_bench_time.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2ar_bench_ref.json 2>/dev/null; tail -c 200 gpurun_out/r2ar_bench_ref.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ar_bench_c4.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fp64_pipe_frac"], d["gpu_launches"])
print({k: round(v["kernel_ms"],1) for k,v in d["per_family"].items()})
for k,v in d["secondary"].items():
    if isinstance(v,dict): print(k, v.get("value"), v.get("unit"), v.get("ms_per_step"))
PY
