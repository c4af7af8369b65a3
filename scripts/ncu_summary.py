"""Condense `ncu --page raw --csv` (+ optional `--page source --csv`) exports into the short text summaries kept under
profiles/.   python scripts/ncu_summary.py gpurun_out/<stem> > profiles/<name>.txt"""
import collections
import csv
import re
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
STALL = "smsp__average_warps_issue_stalled_"


def main(stem):
    rows = list(csv.reader(open(stem + ".raw.csv")))
    hdr, units = rows[0], rows[1]
    for rec in rows[2:]:
        d = dict(zip(hdr, rec))
        print("kernel:", d.get("Kernel Name"), " grid", d.get("Grid Size"), " block", d.get("Block Size"))
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"  {k:75s} {d[k]:>18s} {units[hdr.index(k)]}")
        st = sorted(((float(d[k]), k[len(STALL):-len('_per_issue_active.ratio')]) for k in hdr if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and d[k]), reverse=True)
        print("  warp stall cycles per issued instruction:", ", ".join(f"{n} {v:.2f}" for v, n in st[:7]))
    try:
        rows = list(csv.reader(open(stem + ".source.csv")))
    except OSError:
        return
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    inst = collections.Counter(); samp = collections.Counter(); tot = 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        try:
            n = int(r[ix["# Samples"]]); ie = int(r[ix["Instructions Executed"]])
        except ValueError:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        op = ".".join(m.group(2).split(".")[:2]) if m else "?"
        inst[op] += ie; samp[op] += n; tot += n
    ti = sum(inst.values())
    print(f"  SASS: {len(rows) - 2} instructions; executed warp instructions {ti:.4g}; top opcodes (share of executed / share of stall samples):")
    print("   ", ", ".join(f"{k} {100 * v / ti:.1f}%/{100 * samp[k] / max(tot, 1):.1f}%" for k, v in inst.most_common(14)))


if __name__ == "__main__":
    main(sys.argv[1])
