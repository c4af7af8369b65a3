#!/bin/bash
# round 2, session Q: converged-box retirement (parity with its oracle), C5 at 1e-8 with it, launch list of the c5 bench step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rounds_batch.py tests/test_gpu_two_level.py -x -q -m gpu -k "not c4_stated" > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2q_tests.log
timeout 300 python scripts/probe_c5_two_level.py 1e-8 > gpurun_out/r2q_c5_1e-8.log 2>&1; echo "c5 rc=$?"; cut -c1-1100 gpurun_out/r2q_c5_1e-8.log
timeout 300 python bench.py --workload c5 --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/r2q_bench_c5.json 2> gpurun_out/r2q_bench_c5.err; echo "bench c5 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2q_launches_c5.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/r2q_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open("gpurun_out/r2q_launches_c5.csv")))
hdr=None; agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if "Kernel Name" in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); 
        try: v=float(d["Metric Value"].replace(",",""))
        except: continue
        u=d["Metric Unit"]; v*= {"ns":1e-6,"us":1e-3,"ms":1.0,"s":1e3}.get(u,1e-6)
        k=d["Kernel Name"][:60]; agg[k][0]+=1; agg[k][1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:12]: print(f"{t:10.3f} ms {n:6d}  {k}")
PY
