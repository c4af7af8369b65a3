"""Executed FP64 work of a kernel from an `ncu --set full` capture (where a kernel strength-reduces, the roofline fraction is
reported from EXECUTED FP64 instructions, never from the nominal count).

    python scripts/executed_fp64.py <stem>.raw.csv <work json written by scripts/profile_kernel.py> <key> [profiles/r02_executed_fp64.json]

Reads smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on.sum (thread-level instruction counts of the captured launch; from a
`--set full` capture: the .sum.per_cycle_elapsed rate x smsp__cycles_elapsed.avg), divides by
the launch's integrand evaluations and stores the per-evaluation counts under <key> (e.g. "hcub_rb/genz_product_peak/4") in the tracked JSON
that bench.py multiplies by a run's evaluation count.  flops = 2*dfma + dmul + dadd; pipe slots = dfma + dmul + dadd (one FP64 issue slot each)."""
import csv
import json
import os
import sys


def main(raw, work, key, out="profiles/r02_executed_fp64.json"):
    rows = list(csv.reader(open(raw)))
    hdr = rows[0]
    d = dict(zip(hdr, rows[2]))
    units = dict(zip(hdr, rows[1]))
    w = json.load(open(work))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "ms": 1e6, "us": 1e3, "ns": 1.0, "s": 1e9, "msecond": 1e6, "usecond": 1e3, "nsecond": 1.0, "second": 1e9}
    g = lambda k: float(d[k].replace(",", "")) * scale.get(units.get(k, ""), 1.0)      # bytes and nanoseconds whatever unit ncu chose
    def count(op):
        k = f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum"
        if k in d and d[k] != "":
            return g(k)
        # `--set full` exports the rate: thread instructions per elapsed cycle summed over the SMSPs
        return g(k + ".per_cycle_elapsed") * g("smsp__cycles_elapsed.avg")
    dfma, dmul, dadd = count("dfma"), count("dmul"), count("dadd")
    ev = float(w["evals"])
    rec = {"kernel": d.get("Kernel Name"), "dfma_per_eval": dfma / ev, "dmul_per_eval": dmul / ev, "dadd_per_eval": dadd / ev,
           "flops_per_eval": (2 * dfma + dmul + dadd) / ev, "fp64_inst_per_eval": (dfma + dmul + dadd) / ev,
           "capture": {"evals": w["evals"], "nprob": w.get("nprob"), "gpu_time_ns": g("gpu__time_duration.sum"),
                       "fp64_pipe_pct_ncu": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
                       if "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active" in d else None,
                       "dram_bytes_per_eval": (g("dram__bytes_read.sum") + g("dram__bytes_write.sum")) / ev if "dram__bytes_read.sum" in d else None,
                       "dram_bytes": g("dram__bytes_read.sum") + g("dram__bytes_write.sum") if "dram__bytes_read.sum" in d else None,
                       "source": os.path.basename(raw)}}
    allrec = json.load(open(out)) if os.path.exists(out) else {}
    allrec[key] = rec
    json.dump(allrec, open(out, "w"), indent=1, sort_keys=True)
    print(key, json.dumps(rec))


if __name__ == "__main__":
    main(*sys.argv[1:])
