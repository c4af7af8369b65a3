#!/usr/bin/env python
"""bench.py -- the hot path of integrals.jl_b200 on BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W [--workload c4|c1|c2|c3] [--impl b200|reference]

Default workload (the configuration the metric "fp64 integrals/sec at 1/2/4/8 B200 (batched
GK/Genz-Malik)" is quoted on, BASELINE.json configs[3]): the HCubature Genz suite -- 6 families, 4-D,
2^18 parameter sets per family, reltol 1e-6, abstol 1e-8 (reference default), evaluation cap 4e6
(SURVEY 8(d) C4).  One step = one pass of ib200_hcubature_batch over the 6 x 2^18 problems.
N > 1: one process per GPU (torchrun), every rank solves its own 6 x 2^18 problems (weak scaling, no
data-path collective); `value` = all ranks' integrals / max-over-ranks time.

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, CUDA-event timed.  `e2e`: the same
step through the public solve() interface with host (pinned) inputs and host outputs, copies inside
the timed region.  `roofline`: the dominant kernel against the FP64 DFMA peak measured live (this
path is FP64-pipe bound; the sampled workload c2 is HBM bound against MEASURED_PEAKS.json).
`cpu_baseline` / `--impl reference`: the CPU oracle (oracle/, a C restatement of the reference's
algorithms -- Julia is not available) on all host cores over a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = {"c4": ("fp64 integrals/sec (batched Genz-Malik HCubature, Genz suite 4-D)", "integrals/s"),
          "c1": ("fp64 integrals/sec (batched GK(7,15) QuadGK, sin(x*p) sweep)", "integrals/s"),
          "c3": ("fp64 integrals/sec (fixed tensor Gauss-Legendre n=64 3-D, Genz oscillatory)", "integrals/s"),
          "c2": ("sampled GB/s (trapezoid + Simpson over 4096 x 1048576 fp64)", "GB/s"),
          "c5": ("fp64 Genz-Malik rule applications/sec (one 8-D Genz Gaussian, region-parallel adaptive solve at a fixed box budget)", "boxes/s")}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=list(METRIC))
    ap.add_argument("--nprob", type=int, default=0, help="override the problems per family/batch (smoke runs only)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short c1/c2/c3 side measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--hcub-mode", default="rounds", choices=["rounds", "reference_order"],
                    help="refinement driver of the c4 headline (ib200_hcubature_batch mode): rounds (default) or the reference's one-box-per-step order")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks: nvidia-smi sampled DURING the timed region (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle): cpu_baseline of the b200 arm and the whole --impl reference arm
# ------------------------------------------------------------------------------------------------
def cpu_c4_sample(O, WL, nper, threads):
    """Oracle HCubature (OpenMP over problems) on the first `nper` problems of every family."""
    c = WL.C4
    t0 = time.perf_counter()
    for fam in WL.GENZ:
        par = WL.genz_params(fam, c["dim"], nper)
        O.hcubature_batch(FAM_ORC[fam], np.zeros(c["dim"]), np.ones(c["dim"]), par, reltol=c["reltol"], abstol=c["abstol"],
                          maxiters=c["maxiters"], initdiv=c["initdiv"], nthreads=threads)
    return 6 * nper, time.perf_counter() - t0


def cpu_c1_sample(O, WL, nprob, threads):
    c = WL.C1
    par = WL.c1_params(c["nprob"])[:, :nprob]
    t0 = time.perf_counter()
    O.quadgk_batch("sinxp", c["lb"], c["ub"], par, reltol=c["reltol"], abstol=c["abstol"], nthreads=threads)
    return nprob, time.perf_counter() - t0


def cpu_c3_sample(O, WL, nprob, threads):
    c = WL.C3
    x, w = np.polynomial.legendre.leggauss(c["n1d"])
    par = WL.genz_params(c["family"], c["dim"], nprob)
    t0 = time.perf_counter()
    O.fixed_tensor_batch(FAM_ORC[c["family"]], np.zeros(3), np.ones(3), par, x, w, nthreads=threads)
    return nprob, time.perf_counter() - t0


def cpu_c2_sample(O, WL, n, threads):
    m = WL.C2["m"]
    rng = np.random.default_rng(WL.SEED)
    y = np.asfortranarray(rng.random((m, n)) + 1.0)
    t0 = time.perf_counter()
    for rule in ("trapezoid", "simpson"):
        O.sampled_evalrule(rule, y, h=1.0 / (n - 1), dim=1, nthreads=threads)
    dt = time.perf_counter() - t0
    return 2 * (8.0 * m * n + 8 * m) / 1e9, dt        # GB


def executed_table():
    """profiles/r02_executed_fp64.json: executed FP64 instructions per integrand evaluation of every hot kernel, from the committed
    `ncu --set full` captures (scripts/executed_fp64.py).  `roofline.achieved` = these counts x the run's evaluations / kernel time."""
    path = os.path.join(ROOT, "profiles", "r02_executed_fp64.json")
    return json.load(open(path)) if os.path.exists(path) else {}


def host_cores():
    """the cores this process may run on (NOT OMP_NUM_THREADS: torchrun exports OMP_NUM_THREADS=1 to every rank)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


FAM_ORC = {"sinxp": "sinxp", "genz_oscillatory": "osc", "genz_product_peak": "prodpeak", "genz_corner_peak": "corner",
           "genz_gaussian": "gauss", "genz_c0": "c0", "genz_discontinuous": "discont"}


def cpu_leg(workload, scale=1.0):
    """-> (units, seconds, cores, sample description).  Sized for ~10-30 s of CPU work at scale 1."""
    import oracle as O
    import integrals_b200.workloads as WL
    O.use_baseline_build()                     # -O3 -march=native build of the same source, compiled on this machine (oracle/Makefile)
    cores = host_cores()
    if workload == "c4":
        nper = max(8, int(32 * cores * scale))
        units, dt = cpu_c4_sample(O, WL, nper, cores)
        return units, dt, cores, f"first {nper} problems of each of the 6 families (same seeds, tolerances and 4e6 cap)"
    if workload == "c1":
        n = WL.C1["nprob"]
        units, dt = cpu_c1_sample(O, WL, n, cores)
        return units, dt, cores, f"all {n} problems"
    if workload == "c3":
        n = max(8, int(256 * cores * scale))
        units, dt = cpu_c3_sample(O, WL, n, cores)
        return units, dt, cores, f"first {n} problems (262,144 nodes each)"
    if workload == "c5":
        lb, ub, par, _ = WL.c5_problem()
        nb = max(1 << 12, int((1 << 17) * scale))
        t0 = time.perf_counter()
        _, _, ne, _ = O.hcubature_batch("gauss", lb, ub, par.reshape(-1, 1), reltol=WL.C5["reltol"], abstol=WL.C5["abstol"], maxiters=nb * 401, nthreads=1)
        return int(ne[0]) // 401, time.perf_counter() - t0, 1, (f"the same integral at a budget of {nb} boxes, reference order (one box at a time, "
                                                                 "heap): a single HCubature solve is sequential, so 1 thread")
    n = 1 << 15
    units, dt = cpu_c2_sample(O, WL, n, cores)
    return units, dt, cores, f"4096 x {n} fp64 (1 GiB), trapezoid + Simpson"


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  Julia is not in this image, so this is
    the oracle port (C, OpenMP over problems, all host threads), each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    metric, unit = METRIC[args.workload]
    for _ in range(args.warmup):
        cpu_leg(args.workload, 0.25)
    tot_u = tot_t = 0.0
    for _ in range(args.steps):
        u, dt, cores, sample = cpu_leg(args.workload, 0.25)
        tot_u += u; tot_t += dt
    v = tot_u / tot_t
    import oracle as O
    print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
                      "scaling": "strong" if args.workload == "c5" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": workload_config(args.workload, args, sample=sample),
                      "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample, "build": O.BASELINE_FLAGS},
                      "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(w, args, **extra):
    import integrals_b200.workloads as WL
    if w == "c4":
        c = dict(WL.C4)
        cfg = {"workload": "C4: HCubature Genz suite, 6 families x 2^18 parameter sets, 4-D, [0,1]^4",
               "families": WL.GENZ, "reltol": c["reltol"], "abstol": c["abstol"], "maxiters_evals": c["maxiters"],
               "nprob_per_family_per_gpu": args.nprob or c["nprob_per_family"], "parallelism": f"problems sharded x{args.gpus}, no collective"}
    elif w == "c1":
        cfg = {"workload": "C1: QuadGK int_-2^5 sin(x p) dx, 65,536 p in [0.1, 64]", "reltol": 1e-10, "abstol": 1e-8,
               "nprob_per_gpu": args.nprob or WL.C1["nprob"]}
    elif w == "c3":
        cfg = {"workload": "C3: tensor Gauss-Legendre n=64 in 3-D, 2^20 Genz oscillatory parameter sets",
               "nprob_per_gpu": args.nprob or WL.C3["nprob"]}
    elif w == "c5":
        cfg = {"workload": "C5: one 8-D Genz Gaussian, axes 1-2 on (-inf, inf), axes 3-8 on [0,1], reltol 1e-9, abstol 0, fixed box budget",
               "log2_box_budget": args.nprob or 26, "parallelism": f"one domain, evaluation partitioned x{args.gpus} over a replicated box store, results exchanged through peer inboxes (NVLink), one 8-double NCCL all-gather per round"}
    else:
        cfg = {"workload": "C2: TrapezoidalRule + SimpsonsRule on 4096 x 1,048,576 fp64 along dim 2 (32 GiB, > L2)",
               "n_per_gpu": args.nprob or WL.C2["n"]}
    cfg["l2"] = "256 MB buffer written between timed steps (L2 flush)"
    cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import integrals_b200 as ib
    import integrals_b200.workloads as WL

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference for the CPU leg)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = ib.Engine(local)
    if world > 1:
        eng.comm_init_from_torch()                       # the library's own NCCL communicator (one rank per GPU): single-domain solves
    # one explicit stream for torch and the engine, so torch.cuda.Event brackets the engine's kernels
    # (the legacy default stream has handle 0, which ib200_set_stream reads as "the context's own stream")
    bench_stream = torch.cuda.Stream()
    torch.cuda.set_stream(bench_stream)
    assert bench_stream.cuda_stream != 0
    eng.set_stream(bench_stream.cuda_stream)
    F = ib.FAMILY_IDS
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fp64_peak, _ = eng.fp64_peak()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxrank(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # -------------------------------------------------------------------------------- workloads
    def make_c4(scale=1.0, hcub_mode=None):
        c = WL.C4; d = c["dim"]
        nper = args.nprob or c["nprob_per_family"]
        host, dev = {}, {}
        for fam in WL.GENZ:
            par = WL.genz_params(fam, d, nper, seed=WL.SEED + 1000 * rank, scale=scale)
            pinned = torch.from_numpy(np.ascontiguousarray(par.T)).pin_memory()        # [nprob, 2d] == 2d x nprob col-major
            host[fam] = pinned
            dev[fam] = dict(par=pinned.cuda(), I=torch.empty(nper, dtype=torch.float64, device="cuda"),
                            E=torch.empty(nper, dtype=torch.float64, device="cuda"),
                            N=torch.empty(nper, dtype=torch.int64, device="cuda"),
                            S=torch.empty(nper, dtype=torch.int32, device="cuda"))
        lb, ub = np.zeros(d), np.ones(d)
        kms = {fam: 0.0 for fam in WL.GENZ}
        mode = hcub_mode or args.hcub_mode
        kname, kkey = ("hb::hcub_rb_kernel", "hcub_rb") if mode == "rounds" else ("hp::hcub_pair_kernel", "hcub_pair")
        table = executed_table()

        def step(record=False):
            for fam in WL.GENZ:
                b = dev[fam]
                eng.hcubature_batch(F[fam], d, lb, ub, b["par"], reltol=c["reltol"], abstol=c["abstol"], maxevals=c["maxiters"],
                                    initdiv=c["initdiv"], out_I=b["I"], out_E=b["E"], out_nevals=b["N"], out_status=b["S"], mode=mode)
                if record:
                    kms[fam] += eng.last_kernel_ms

        def e2e_step():
            res = 0.0
            for fam in WL.GENZ:
                prob = ib.IntegralProblem(getattr(ib, ib_marker[fam])(), (lb, ub), host[fam].numpy().T)
                sol = ib.solve(prob, ib.HCubatureB200(initdiv=c["initdiv"], refinement=mode), engine=eng, reltol=c["reltol"], abstol=c["abstol"],
                               maxiters=c["maxiters"])
                res += float(sol.u[0])
            return res

        def finish(kernel_ms_total):
            fl = 0.0; ev = 0; per = {}; flx = 0.0; have_all = True
            for fam in WL.GENZ:
                n_ev = int(dev[fam]["N"].sum().item())
                ev += n_ev
                fl += n_ev * WL.hcubature_flops_per_eval(fam, d)
                ex = table.get(f"{kkey}/{fam}/{d}")
                have_all = have_all and ex is not None
                flx += n_ev * (ex["flops_per_eval"] if ex else 0.0)
                ms = kms[fam] / max(args.steps, 1)
                per[fam] = {"kernel_ms": ms, "integrals_per_s": nper / (ms * 1e-3) if ms else None,
                            "mean_evals": n_ev / nper, "frac_maxevals": float((dev[fam]["S"] == 1).double().mean().item()),
                            "gevals_per_s": n_ev / (ms * 1e-3) / 1e9 if ms else None,
                            "executed_fp64_tflops": n_ev * ex["flops_per_eval"] / (ms * 1e-3) / 1e12 if (ex and ms) else None,
                            "executed_fp64_inst_per_eval": ex["fp64_inst_per_eval"] if ex else None}
            dom = max(WL.GENZ, key=lambda f: kms[f])
            dom_ev = int(dev[dom]["N"].sum().item())
            return dict(units=6 * nper, launches=6, flops=fl, flops_executed=flx if have_all else None, evals=ev, per_family=per,
                        h2d=6 * nper * 2 * d * 8 + 6 * 2 * d * 8, d2h=6 * nper * (8 + 8 + 8 + 4),
                        kernel=f"{kname}<FAM,4,true> (one launch per family and step" + (
                            "; rounds mode: compact 32-byte box records, and a full-record launch behind each that picks up handed-back problems)"
                            if mode == "rounds" else ")"), refinement=mode,
                        dominant=dict(kernel=f"{kname}<{dom},4,true>", key=f"{kkey}/{dom}/{d}", ms=kms[dom] / max(args.steps, 1), evals=dom_ev,
                                      flops=dom_ev * WL.hcubature_flops_per_eval(dom, d),
                                      share_of_step=kms[dom] / max(sum(kms.values()), 1e-9), nprob=nper,
                                      default_size=(nper == c["nprob_per_family"])))
        return step, e2e_step, finish

    ib_marker = {"genz_oscillatory": "GenzOscillatory", "genz_product_peak": "GenzProductPeak", "genz_corner_peak": "GenzCornerPeak",
                 "genz_gaussian": "GenzGaussian", "genz_c0": "GenzC0", "genz_discontinuous": "GenzDiscontinuous"}

    def make_c1_vec(ncomp=4):
        """the C1 sweep as 65,536/ncomp vector-valued problems of ncomp components each (shared subdivision, 2-norm)"""
        c = WL.C1
        n = (args.nprob or c["nprob"]) // ncomp
        par = np.asfortranarray(WL.c1_params(c["nprob"])[:, :n * ncomp].reshape(1, ncomp, n, order="F"))
        last = {}

        def step(record=False):
            I, E, ne, st = eng.quadgk_vec_batch(F["sinxp"], c["lb"], c["ub"], par, reltol=c["reltol"], abstol=c["abstol"])
            last["ev"] = int(ne.sum())

        def finish(kernel_ms_total):
            return dict(units=n * ncomp, launches=1, flops=last["ev"] * ncomp * WL.quadgk_flops_per_eval(), evals=last["ev"] * ncomp, h2d=n * ncomp * 8, d2h=n * (ncomp * 8 + 20),
                        kernel="quadgk_vec_kernel<SINXP,true,4> (host buffers: the vector entry point takes host arrays)")
        return step, step, finish

    def make_c1(abstol=None, order=7):
        c = dict(WL.C1)
        if abstol is not None:
            c["abstol"] = abstol
        rule = None
        if order != 7:
            from integrals_b200.kronrod import kronrod
            rule = kronrod(order)
        n = args.nprob or c["nprob"]
        par = WL.c1_params(c["nprob"])[:, :n]
        pinned = torch.from_numpy(np.ascontiguousarray(par.T)).pin_memory()
        dpar = pinned.cuda()
        oI = torch.empty(n, dtype=torch.float64, device="cuda"); oE = torch.empty_like(oI)
        oN = torch.empty(n, dtype=torch.int64, device="cuda"); oS = torch.empty(n, dtype=torch.int32, device="cuda")
        lb, ub = np.array([c["lb"]]), np.array([c["ub"]])

        def step(record=False):
            eng.quadgk_batch(F["sinxp"], lb, ub, dpar, reltol=c["reltol"], abstol=c["abstol"], out_I=oI, out_E=oE, out_nevals=oN, out_status=oS,
                             rule=rule)

        def e2e_step():
            sol = ib.solve(ib.IntegralProblem(ib.SinXP(), (c["lb"], c["ub"]), pinned.numpy().T), ib.QuadGKB200(), engine=eng,
                           reltol=c["reltol"])
            return float(sol.u[0])

        def finish(kernel_ms_total):
            ev = int(oN.sum().item())
            return dict(units=n, launches=1, flops=ev * WL.quadgk_flops_per_eval(), evals=ev, h2d=n * 8 + 16, d2h=n * 28,
                        kernel="quadgk_kernel<SINXP,true>", key=f"quadgk/sinxp/order{order}", nprob=n)
        return step, e2e_step, finish

    def make_c3():
        c = WL.C3
        n = args.nprob or c["nprob"]
        x, w = np.polynomial.legendre.leggauss(c["n1d"])
        par = WL.genz_params(c["family"], c["dim"], n, seed=WL.SEED + 1000 * rank)
        pinned = torch.from_numpy(np.ascontiguousarray(par.T)).pin_memory()
        dpar = pinned.cuda()
        out = torch.empty(n, dtype=torch.float64, device="cuda")
        lb, ub = np.zeros(3), np.ones(3)

        def step(record=False):
            eng.fixed_rule_tensor_batch(F[c["family"]], 3, lb, ub, dpar, x, w, out=out)

        def e2e_step():
            sol = ib.solve(ib.IntegralProblem(ib.GenzOscillatory(), (lb, ub), pinned.numpy().T),
                           ib.QuadratureRuleB200(lambda k: (x, w), n=c["n1d"], tensor=True), engine=eng)
            return float(sol.u[0])

        def finish(kernel_ms_total):
            ev = n * c["n1d"] ** 3
            return dict(units=n, launches=1, flops=ev * WL.fixed_tensor_flops_per_eval(c["family"], 3), evals=ev,
                        h2d=n * 6 * 8, d2h=n * 8, kernel="tensor_kernel<GENZ_OSC,3>", key="tensor/genz_oscillatory/3", nprob=n)
        return step, e2e_step, finish

    def make_c2(nonuniform=False, contiguous=False):
        """default: the integration axis is the strided one (dim = 2 of a column-major 4096 x n array), uniform grid.
        nonuniform: x_i = ((i-1)/(n-1))^1.5; contiguous: the transposed array, dim = 1 (inner = 1)."""
        m = WL.C2["m"]; n = args.nprob or WL.C2["n"]
        xs = torch.linspace(0, 1, n, dtype=torch.float64, device="cuda")
        jj = 1.0 + torch.arange(1, m + 1, dtype=torch.float64, device="cuda") / m
        if contiguous:
            y = torch.empty((m, n), dtype=torch.float64, device="cuda")     # row-major [m, n] == column-major n x m, dim = 1
            for j0 in range(0, m, 64):
                y[j0:j0 + 64] = torch.sin(jj[j0:j0 + 64, None] * xs[None, :]) + 1.5
        else:
            y = torch.empty((n, m), dtype=torch.float64, device="cuda")     # row-major [n, m] == column-major m x n, dim = 2
            for i0 in range(0, n, 1 << 16):
                i1 = min(n, i0 + (1 << 16))
                y[i0:i1] = torch.sin(xs[i0:i1, None] * jj[None, :]) + 1.5
        out = torch.empty(m, dtype=torch.float64, device="cuda")
        h = 1.0 / (n - 1)
        xg = xs ** 1.5 if nonuniform else None
        inner, outer = (1, m) if contiguous else (m, 1)
        ne2e = min(n, 1 << 16)                                              # e2e on a 2 GiB host block (PCIe bound)
        yh = None if (nonuniform or contiguous) else y[:ne2e].cpu().pin_memory()

        def step(record=False):
            eng.sampled("trapezoid", y, inner, n, outer, h=h, x=xg, out=out)
            eng.sampled("simpson", y, inner, n, outer, h=h, x=xg, out=out)

        def e2e_step():
            r = 0.0
            for alg in (ib.TrapezoidalB200(), ib.SimpsonsB200()):
                sol = ib.solve(ib.SampledIntegralProblem(yh.numpy().T, ib.Range(0.0, h * (ne2e - 1), ne2e), dim=1), alg, engine=eng)
                r += float(sol.u[0])
            return r

        def finish(kernel_ms_total):
            by = 2 * (8.0 * m * n + 8 * m)
            return dict(units=by / 1e9, launches=6, bytes=by, h2d=2 * 8 * m * ne2e, d2h=2 * 8 * m, e2e_units=2 * (8.0 * m * ne2e + 8 * m) / 1e9,
                        kernel="strided_kernel<2,8> (2 launches per step: trapezoid, Simpson; + weights and finish kernels)")
        return step, e2e_step, finish

    def make_c2_f32():
        """Float32 samples on a Float32 grid (ib200_sampled_f32), 4096 x 2^20, strided axis"""
        m = WL.C2["m"]; n = args.nprob or WL.C2["n"]
        xs = torch.linspace(0, 1, n, dtype=torch.float32, device="cuda")
        jj = 1.0 + torch.arange(1, m + 1, dtype=torch.float32, device="cuda") / m
        y = torch.empty((n, m), dtype=torch.float32, device="cuda")
        for i0 in range(0, n, 1 << 16):
            i1 = min(n, i0 + (1 << 16))
            y[i0:i1] = torch.sin(xs[i0:i1, None] * jj[None, :]) + 1.5
        out = torch.empty(m, dtype=torch.float32, device="cuda")
        h = 1.0 / (n - 1)

        def step(record=False):
            eng.sampled_f32("trapezoid", y, m, n, 1, h=h, out=out)
            eng.sampled_f32("simpson", y, m, n, 1, h=h, out=out)

        def finish(kernel_ms_total):
            by = 2 * (4.0 * m * n + 4 * m)
            return dict(units=by / 1e9, launches=6, bytes=by, h2d=0, d2h=2 * 4 * m, kernel="strided_kernel<float,2,8>")
        return step, step, finish

    def make_c5(reltol=None):
        """reltol None: the stated tolerance (1e-9) at a fixed budget of rule applications, single level (status 1: a throughput figure).
        reltol given: the two-level solve run TO that tolerance (status 0: a time-to-tolerance figure)."""
        lb, ub, par, exact = WL.c5_problem()
        to_tol = reltol is not None
        budget = (1 << 40) if to_tol else 1 << (args.nprob or 26)                # --nprob overrides log2 of the box budget
        rt = reltol if to_tol else WL.C5["reltol"]
        last = {}

        def step(record=False):
            eng.set_option("single_switch_boxes", (1 << 20) if to_tol else 0)
            I, E, st = eng.hcubature_single(F["genz_gaussian"], 8, lb, ub, par, reltol=rt, abstol=WL.C5["abstol"], max_boxes=budget)
            eng.set_option("single_switch_boxes", 1 << 20)
            last.update(I=I, E=E, **st)

        def e2e_step():
            eng.set_option("single_switch_boxes", (1 << 20) if to_tol else 0)
            sol = ib.solve(ib.IntegralProblem(ib.GenzGaussian(), (lb, ub), par), ib.HCubatureB200(single_domain=True), engine=eng,
                           reltol=rt, abstol=WL.C5["abstol"], maxiters=min(budget * 401, 1 << 62))
            eng.set_option("single_switch_boxes", 1 << 20)
            return float(sol.u)

        def finish(kernel_ms_total):
            ev = last["nevals"]
            return dict(units=last["rule_applications"], global_units=True, launches=8 * last["rounds"] + 4, evals=ev,
                        flops=ev * (WL.FAMILY_FLOPS["genz_gaussian"](8) + 2 * 20 + 6 * 4 + 2 * 8 + 1 + 30.0 / 401), h2d=8 * 32, d2h=8 * 8,
                        kernel="hs::eval_thread_kernel<genz_gaussian,8,false> (+ 7 bookkeeping kernels per refinement round; two-level solves: hb::hcub_rb_kernel<genz_gaussian,8,false>)", key="hcub_rb/genz_gaussian/8",
                        result=dict(I=last["I"], E=last["E"], rel_E=last["E"] / abs(last["I"]), true_rel_err=abs(last["I"] - exact) / abs(exact),
                                    reltol=rt, rounds=last["rounds"], boxes=last["boxes"], rule_applications=last["rule_applications"],
                                    level2_subproblems=last.get("level2_subproblems", 0), status=last["status"],
                                    box_budget=None if to_tol else budget))
        return step, e2e_step, finish

    def make_c4_vec(ncomp=4, dim=3):
        """vector-valued HCubature -- 16,384 problems of 4 Genz Gaussian components each in 3-D (the components
        share points and subdivision, error = 2-norm over them), reduced difficulty (h/4)"""
        n = args.nprob or 16384
        par = np.asfortranarray(WL.genz_params("genz_gaussian", dim, n * ncomp, scale=0.25).reshape(2 * dim, ncomp, n, order="F"))
        lb, ub = np.zeros(dim), np.ones(dim)
        last = {}

        def step(record=False):
            I, E, ne, st = eng.hcubature_vec_batch(F["genz_gaussian"], dim, lb, ub, par, reltol=1e-6, abstol=1e-8, maxevals=400000)
            last.update(ev=int(ne.sum()), capped=float((st != 0).mean()))

        def finish(kernel_ms_total):
            per_point = ncomp * (WL.FAMILY_FLOPS["genz_gaussian"](dim) + 1) + WL.TRANSFORM_FLOPS_FINITE * dim + 2 * dim
            return dict(units=n * ncomp, launches=1, flops=last["ev"] * per_point, evals=last["ev"] * ncomp, h2d=par.nbytes, d2h=n * (ncomp * 8 + 20),
                        kernel="hv::hcub_vec_kernel<genz_gaussian,3> (host buffers: the vector entry point takes host arrays)",
                        extra=dict(problems=n, components=ncomp, points=last["ev"], frac_maxevals=last["capped"]))
        return step, step, finish

    def make_bridge(dim=4):
        """one 4-D Genz Gaussian through the BatchIntegralFunction bridge -- the engine hands out each round's
        points on the device, the integrand is a torch expression evaluated by the caller on the same buffers"""
        par = WL.genz_params("genz_gaussian", dim, 1)[:, 0]
        a = torch.tensor(par[:dim], dtype=torch.float64, device="cuda"); u = torch.tensor(par[dim:], dtype=torch.float64, device="cuda")
        lb, ub = np.zeros(dim), np.ones(dim)
        budget = 1 << (args.nprob or 18)
        last = {}

        def f(x):
            t = a * (x - u)
            return torch.exp(-(t * t).sum(dim=1))

        def step(record=False):
            I, E, st = eng.integrate_batch(f, dim, lb, ub, reltol=1e-9, abstol=0.0, max_boxes=budget, device=True)
            last.update(I=I, E=E, **st)

        def finish(kernel_ms_total):
            ev = last["nevals"]
            return dict(units=ev, global_units=True, launches=9 * last["batches"] + 3, flops=0.0, evals=ev, h2d=0, d2h=0,
                        bytes=ev * (8 * (dim + 1) + 16),     # points_kernel writes D + 1 doubles per point, combine_kernel reads 2
                        kernel="bb::points_kernel<4> / bb::combine_kernel<4> + 7 bookkeeping kernels per round; integrand: torch on the same stream",
                        extra=dict(I=last["I"], rel_E=last["E"] / abs(last["I"]), rounds=last["rounds"], batches=last["batches"],
                                   boxes=last["boxes"], status=last["status"], box_budget=budget))
        return step, step, finish

    makers = {"c4": make_c4, "c1": make_c1, "c2": make_c2, "c3": make_c3, "c5": make_c5,
              "c5_to_tolerance_1e-7": lambda: make_c5(reltol=1e-7),
              "c4_reference_order": lambda: make_c4(hcub_mode="reference_order"),
              # side measurements SURVEY 8(d) asks for next to the configs (reported under "secondary" only)
              "c1_abstol_1e-12": lambda: make_c1(abstol=1e-12), "c2_nonuniform_grid": lambda: make_c2(nonuniform=True),
              "c2_contiguous_axis": lambda: make_c2(contiguous=True), "c4_reduced_difficulty_h/4": lambda: make_c4(scale=0.25),
              # "next" rows of SURVEY 8(f), same measurement as their parent config
              "c1_order_15": lambda: make_c1(order=15), "c1_vector_valued_4": make_c1_vec, "c2_float32": make_c2_f32,
              "hcubature_vector_valued_3d_4": make_c4_vec, "batch_bridge_device_integrand_4d": make_bridge}
    EXTRA_UNITS = {"c5_to_tolerance_1e-7": "boxes/s", "c4_reference_order": "integrals/s", "c1_abstol_1e-12": "integrals/s", "c2_nonuniform_grid": "GB/s", "c2_contiguous_axis": "GB/s",
                   "c4_reduced_difficulty_h/4": "integrals/s", "c1_order_15": "integrals/s", "c1_vector_valued_4": "integrals/s",
                   "c2_float32": "GB/s", "hcubature_vector_valued_3d_4": "integrals/s", "batch_bridge_device_integrand_4d": "points/s"}

    def run(workload, steps, warmup, with_e2e=True, sample_clocks=False):
        step, e2e_step, finish = makers[workload]()
        tw = time.perf_counter()
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        while world == 1 and time.perf_counter() - tw < 0.3:   # millisecond-scale workloads: keep warming until the clocks have ramped
            # (single rank only: a time-based count could differ between ranks, and the c5 step contains a collective)
            step()
            torch.cuda.synchronize()
        barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        l0 = eng.kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            flush.fill_(1)                                  # L2 flush (inside the timed region; ~0.05 ms)
            step(record=True)
        e1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = maxrank(e0.elapsed_time(e1))
        launches = eng.kernel_launches - l0
        info = finish(ms)
        mult = 1 if info.get("global_units") else world      # c5: one domain over all ranks (units are already global)
        res = dict(ms_per_step=ms / steps, value=mult * info["units"] / (ms / steps * 1e-3), launches=launches, info=info, clocks=clocks)
        if with_e2e:
            e2e_step()                                      # warm-up (pinned staging buffers, scratch growth)
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                e2e_step()
            torch.cuda.synchronize()
            dt = maxrank(time.perf_counter() - t0)
            res["e2e_value"] = mult * info.get("e2e_units", info["units"]) / (dt / steps)
        return res

    w = args.workload
    metric, unit = METRIC[w]
    r = run(w, args.steps, args.warmup, with_e2e=True, sample_clocks=True)
    info = r["info"]
    # roofline of the dominant kernel: algorithmic flops (or bytes) per launch / mean launch duration (CUDA events)
    if w == "c2":
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        peak = peaks.get("hbm_gbs", 6650.0)
        ach = info["bytes"] / (r["ms_per_step"] * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6.65 TB/s (of fallback)", "kernel": info["kernel"]}
    else:
        # FP64 roofline from EXECUTED instructions: per-evaluation DFMA/DMUL/DADD counts of the committed ncu capture of
        # this kernel (profiles/r02_executed_fp64.json) x the evaluations of this run / the kernel's CUDA-event time.  The nominal
        # (algorithmic) count of SURVEY 8(d) is kept beside it: the kernels execute fewer operations than it assumes.
        table = executed_table()
        dom = info.get("dominant")
        if dom:                         # the dominant kernel of the step: one launch, CUDA-event timed on the launching stream
            kname, kev, kms_, nom_flops, key = dom["kernel"], dom["evals"], dom["ms"], dom["flops"], dom.get("key")
        else:
            kname, kev, kms_, nom_flops, key = info["kernel"], info["evals"], r["ms_per_step"], info["flops"], info.get("key")
        ex = table.get(key) if key else None
        sec_k = kms_ * 1e-3
        nominal = nom_flops / sec_k / 1e12
        ach = kev * ex["flops_per_eval"] / sec_k / 1e12 if ex else None
        traffic = traffic_note = None
        if ex and ex["capture"].get("dram_bytes_per_eval") is not None:
            same = ex["capture"].get("nprob") == (dom or {}).get("nprob", info.get("nprob"))
            traffic = ex["capture"]["dram_bytes"] if same else ex["capture"]["dram_bytes_per_eval"] * kev
            traffic_note = (f"profiles/{ex['capture']['source']}: dram__bytes_read.sum + dram__bytes_write.sum of one launch"
                            + ("" if same else f" of {ex['capture'].get('nprob')} problems, scaled by evaluations to this launch"))
        roof = {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak if ach is not None else None,
                "traffic": traffic, "traffic_source": traffic_note,
                "flops_model": "EXECUTED FP64: (2 DFMA + DMUL + DADD thread instructions per integrand evaluation, from the ncu capture named in "
                               "executed_source) x evaluations of this launch / CUDA-event kernel time" if ex else "executed counts missing for this kernel",
                "executed_source": f"profiles/r02_executed_fp64.json[{key}] <- profiles/{ex['capture']['source']}" if ex else None,
                "executed_flops_per_eval": ex["flops_per_eval"] if ex else None,
                "fp64_pipe_frac": kev * ex["fp64_inst_per_eval"] / sec_k / (fp64_peak * 1e12 / 2) if ex else None,
                "fp64_pipe_frac_note": "FP64 issue slots used (DFMA, DMUL and DADD take one slot each; peak = measured DFMA rate)",
                "fp64_pipe_pct_in_capture_ncu": ex["capture"].get("fp64_pipe_pct_ncu") if ex else None,
                "nominal_achieved": nominal, "nominal_frac": nominal / fp64_peak,
                "nominal_note": "SURVEY 8(d) algorithmic operation count per evaluation x evaluations (exceeds the executed count: separable-term tables, angle addition)",
                "peak_source": "builder-measured: DFMA-chain microbenchmark run live on this GPU (ib200_measure_fp64_peak; nominal 37.2); MEASURED_PEAKS.json has no FP64 entry",
                "kernel": kname, "kernel_ms": kms_, "evals_per_launch": kev, "gevals_per_s": kev / sec_k / 1e9,
                "share_of_step": dom["share_of_step"] if dom else 1.0}
        if info.get("flops_executed"):
            roof["whole_step"] = {"achieved": info["flops_executed"] / (r["ms_per_step"] * 1e-3) / 1e12,
                                  "frac": info["flops_executed"] / (r["ms_per_step"] * 1e-3) / 1e12 / fp64_peak, "evals": info["evals"],
                                  "gevals_per_s": info["evals"] / (r["ms_per_step"] * 1e-3) / 1e9}
    line = {"metric": metric, "value": r["value"], "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong" if w == "c5" else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(w, args), "clocks": r["clocks"],
            "e2e": {"value": r["e2e_value"], "unit": unit, "h2d_bytes_per_step": info["h2d"], "d2h_bytes_per_step": info["d2h"],
                    "api": "integrals_b200.solve(IntegralProblem | SampledIntegralProblem, alg) with host numpy (pinned) inputs and host outputs"},
            "gpu_launches": r["launches"], "roofline": roof}
    if "per_family" in info:
        line["per_family"] = info["per_family"]
    if "result" in info:
        line["result"] = info["result"]

    if rank == 0 and world == 1 and not args.no_secondary:
        sec = {}
        for sw in list(METRIC) + list(EXTRA_UNITS):
            if sw == w:
                continue
            if sw.startswith("c2") and torch.cuda.mem_get_info()[0] < (40 << 30):
                continue
            if sw == "c4_reference_order" and args.hcub_mode == "reference_order":
                continue
            try:
                s = run(sw, 1, 1, with_e2e=False) if sw in ("c4_reference_order", "c5_to_tolerance_1e-7") else run(sw, 2, 3, with_e2e=False)
                e = {"value": s["value"], "unit": METRIC[sw][1] if sw in METRIC else EXTRA_UNITS[sw], "ms_per_step": s["ms_per_step"]}
                if "result" in s["info"]:
                    e["result"] = s["info"]["result"]
                if "per_family" in s["info"]:
                    e["frac_maxevals"] = {k: v["frac_maxevals"] for k, v in s["info"]["per_family"].items()}
                if "extra" in s["info"]:
                    e.update(s["info"]["extra"])
                if sw.startswith("c2"):
                    e["frac_of_measured_hbm"] = s["value"] / 6535.7
                elif "bytes" in s["info"]:                   # the bridge: HBM bytes of its own two kernels per point
                    e["algorithmic_GBps"] = s["info"]["bytes"] / (s["ms_per_step"] * 1e-3) / 1e9
                else:
                    tab = executed_table()
                    inf = s["info"]
                    ex = tab.get(inf.get("key") or (inf.get("dominant") or {}).get("key"))
                    if inf.get("flops_executed"):              # c4: every family's own count
                        e["fp64_tflops_executed"] = inf["flops_executed"] / (s["ms_per_step"] * 1e-3) / 1e12
                    elif ex:
                        e["fp64_tflops_executed"] = inf["evals"] * ex["flops_per_eval"] / (s["ms_per_step"] * 1e-3) / 1e12
                        e["fp64_pipe_frac"] = inf["evals"] * ex["fp64_inst_per_eval"] / (s["ms_per_step"] * 1e-3) / (fp64_peak * 1e12 / 2)
                    if "fp64_tflops_executed" in e:
                        e["frac_of_fp64_peak"] = e["fp64_tflops_executed"] / fp64_peak
                    e["fp64_tflops_nominal"] = inf["flops"] / (s["ms_per_step"] * 1e-3) / 1e12
                    e["nominal_frac_of_fp64_peak"] = e["fp64_tflops_nominal"] / fp64_peak
                sec[sw] = e
            except Exception as ex:                          # a side measurement must not lose the headline
                sec[sw] = {"error": str(ex)[:200]}
            torch.cuda.empty_cache()
        sec["_note"] = ("frac_of_fp64_peak / fp64_tflops_executed: EXECUTED FP64 instructions (per-evaluation counts of the committed ncu captures, "
                        "profiles/r02_executed_fp64.json) x this run's evaluations / time -- cannot exceed 1.  fp64_tflops_nominal uses the algorithmic "
                        "counts of SURVEY 8(d) and exceeds the peak where a kernel executes fewer operations than the count assumes (angle addition in "
                        "c3, per-axis factor tables in c4 / c5)")
        line["secondary"] = sec

    if world > 1 and not args.no_secondary:
        # every rank takes part; rank 0 reports.  (i) BASELINE config 5 through the library's NCCL communicator -- the single-domain path
        # that the c4 headline never touches; (ii) strong scaling of the FIXED 6 x 2^18 C4 set through sharding.solve_sharded.
        sec = {}

        def bits(*vals):
            t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device="cuda")
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            return all(torch.equal(parts[0].view(torch.int64), q.view(torch.int64)) for q in parts)

        def c5_entry(name, reltol, warm=True, solo_check=True):
            try:
                step, _, finish = makers[name]() if reltol is None else make_c5(reltol=reltol)
                if warm:
                    step()
                barrier()
                t0 = time.perf_counter(); step(); torch.cuda.synchronize()
                dt = maxrank(time.perf_counter() - t0)
                info = finish(0.0); res = info["result"]
                same = bits(res["I"], res["E"], res["rounds"], res["boxes"], res["rule_applications"])
                e = {"ms": dt * 1e3, "value": res["rule_applications"] / dt, "unit": "boxes/s", "result": res, "identical_on_all_ranks": same}
                if rank == 0 and solo_check:                  # the same solve on ONE GPU (a second context without a communicator)
                    solo = ib.Engine(local)
                    solo.set_option("single_switch_boxes", (1 << 20) if reltol is not None else 0)
                    lb5, ub5, par5, _ = WL.c5_problem()
                    t0 = time.perf_counter()
                    I1, E1, s1 = solo.hcubature_single(F["genz_gaussian"], 8, lb5, ub5, par5, reltol=res["reltol"], abstol=WL.C5["abstol"],
                                                       max_boxes=res["box_budget"] or (1 << 40))
                    e["single_gpu_ms_same_box"] = (time.perf_counter() - t0) * 1e3
                    e["equals_single_gpu_bit_for_bit"] = (I1 == res["I"] and E1 == res["E"] and s1["rule_applications"] == res["rule_applications"])
                    solo.close()
                barrier()
                return e
            except Exception as ex:
                return {"error": str(ex)[:300]}
        sec["c5_stated_tolerance_fixed_budget"] = c5_entry("c5", None)
        sec["c5_to_tolerance_1e-7"] = c5_entry("c5_to_tolerance_1e-7", 1e-7)
        if world >= int(os.environ.get("IB200_BENCH_C5_STATED_MIN_WORLD", "8")):
            # BASELINE config 5 as stated: reltol 1e-9 on the 8 GPUs of the box (~1e12 rule applications: minutes on one GPU, so neither a
            # warm-up solve nor the one-GPU comparison).  The two environment variables exist to rehearse this entry on a smaller box.
            rt9 = float(os.environ.get("IB200_BENCH_C5_STATED_RELTOL", "1e-9"))
            sec["c5_stated_config_%g" % rt9] = c5_entry("c5_stated_config", rt9, warm=False, solo_check=False)
        try:
            from integrals_b200.sharding import solve_sharded
            c = WL.C4; d = c["dim"]; nper = args.nprob or c["nprob_per_family"]
            lb4, ub4 = np.zeros(d), np.ones(d)
            pars = {fam: WL.genz_params(fam, d, nper) for fam in WL.GENZ}              # the SAME fixed set on every rank
            def strong():
                tot = 0.0
                for fam in WL.GENZ:
                    prob = ib.IntegralProblem(getattr(ib, ib_marker[fam])(), (lb4, ub4), pars[fam])
                    sol = solve_sharded(prob, ib.HCubatureB200(refinement=args.hcub_mode), interleave=True, engine=eng, reltol=c["reltol"],
                                        abstol=c["abstol"], maxiters=c["maxiters"])
                    tot += float(sol.u.sum())
                return tot
            strong(); barrier()
            t0 = time.perf_counter(); chk = strong(); torch.cuda.synchronize()
            dt = maxrank(time.perf_counter() - t0)
            sec["c4_strong_scaling_fixed_set"] = {"value": 6 * nper / dt, "unit": "integrals/s", "seconds": dt, "problems": 6 * nper,
                                                  "same_sums_on_all_ranks": bits(chk),
                                                  "how": "sharding.solve_sharded(interleave=True): problem k on rank k mod N, host inputs and outputs, one all_gather of the results per family"}
        except Exception as ex:
            sec["c4_strong_scaling_fixed_set"] = {"error": str(ex)[:300]}
        try:
            # BASELINE config 2 with the INTEGRATION axis cut into one slab per GPU: every rank streams its contiguous
            # slab (generated on its device) and the 4096 partial results are summed by ONE NCCL all-reduce inside ib200_sampled_slab
            m = WL.C2["m"]; ntot = args.nprob or WL.C2["n"]
            lo, hi = (ntot * rank) // world, (ntot * (rank + 1)) // world
            xs = torch.linspace(0, 1, ntot, dtype=torch.float64, device="cuda")[lo:hi]
            jj = 1.0 + torch.arange(1, m + 1, dtype=torch.float64, device="cuda") / m
            ys = torch.empty((hi - lo, m), dtype=torch.float64, device="cuda")
            for i0 in range(0, hi - lo, 1 << 16):
                i1 = min(hi - lo, i0 + (1 << 16))
                ys[i0:i1] = torch.sin(xs[i0:i1, None] * jj[None, :]) + 1.5
            out = torch.empty(m, dtype=torch.float64, device="cuda")
            hh = 1.0 / (ntot - 1)
            def split_step():
                eng.sampled_slab("trapezoid", ys, m, hi - lo, 1, ntot, lo, h=hh, out=out)
                eng.sampled_slab("simpson", ys, m, hi - lo, 1, ntot, lo, h=hh, out=out)
            for _ in range(3):
                split_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(4):
                flush.fill_(1); split_step()
            torch.cuda.synchronize()
            dt = maxrank(time.perf_counter() - t0) / 4
            # Simpson of sin(x j') + 1.5 over [0, 1], row j' = 1 + j/m: closed form
            jh = jj.cpu().numpy(); exact = (1 - np.cos(jh)) / jh + 1.5
            errmax = float(np.max(np.abs(out.cpu().numpy() - exact)))
            by = 2 * (8.0 * m * ntot + 8 * m)
            sec["c2_split_integration_axis"] = {"value": by / dt / 1e9, "unit": "GB/s", "ms_per_step": dt * 1e3, "frac_of_measured_hbm_x_gpus": by / dt / 1e9 / (6535.7 * world),
                                                "max_abs_err_vs_closed_form": errmax, "same_result_on_all_ranks": bits(*out[:8].tolist()),
                                                "how": "each rank: ib200_sampled_slab on its 4096 x n/N slab (device resident) + one ncclAllReduce of 4096 doubles per rule"}
            del ys
        except Exception as ex:
            sec["c2_split_integration_axis"] = {"error": str(ex)[:300]}
        line["secondary"] = sec

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        units, dt, cores, sample = cpu_leg(w)
        import oracle as O
        line["cpu_baseline"] = {"value": units / dt, "unit": unit, "cores": cores, "kind": "port", "sample": sample,
                                "seconds": dt, "build": O.BASELINE_FLAGS}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
