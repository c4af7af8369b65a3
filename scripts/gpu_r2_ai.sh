#!/bin/bash
# round 2, session AI: the whole GPU suite, the default bench line, the reference arm, the launch list of one bench step and a full
# capture of the dominant kernel (4-D product peak, compact records) -- after the cold-path and compact-record changes
mkdir -p gpurun_out
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
}
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2ai_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ai_tests.log | cut -c1-300
export PK_MODE=1
PK="python scripts/profile_kernel.py hcub genz_product_peak 32768"
timeout 200 $PK 2>&1 | tail -1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hcub_rb -s 2 -c 1 -o gpurun_out/r02c_rb_genz_product_peak -f $PK > gpurun_out/r2ai_ncu_full.log 2>&1; echo "ncu rc=$?"
export_rep r02c_rb_genz_product_peak
python scripts/executed_fp64.py gpurun_out/r02c_rb_genz_product_peak.raw.csv gpurun_out/pk_hcub_genz_product_peak_32768_m1.json hcub_rb/genz_product_peak/4 profiles/r02_executed_fp64.json | cut -c1-400
cp gpurun_out/r02c_rb_genz_product_peak.raw.csv profiles/; cp profiles/r02_executed_fp64.json gpurun_out/r02_executed_fp64.json
unset PK_MODE
( time timeout 1200 python bench.py > gpurun_out/r2ai_bench_c4.json 2> gpurun_out/r2ai_bench_c4.err ) 2> gpurun_out/r2ai_bench_time.txt; echo "bench rc=$?"; tail -2 gpurun_out/r2ai_bench_c4.err; cat gpurun_out/r2ai_bench_time.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2ai_bench_ref.json 2>/dev/null; tail -c 300 gpurun_out/r2ai_bench_ref.json
timeout 300 python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/r2ai_bench_1step.json 2>/dev/null; echo "1step rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2ai_launches_c4.csv python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/r2ai_ncu_launches.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ai_bench_c4.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fp64_pipe_frac"], d["gpu_launches"])
print({k: round(v["kernel_ms"],1) for k,v in d["per_family"].items()})
for k,v in d["secondary"].items():
    if isinstance(v,dict): print(k, v.get("value"), v.get("unit"), v.get("ms_per_step"))
PY
