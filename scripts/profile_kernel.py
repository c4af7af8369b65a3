"""One short invocation of a hot-path kernel for ncu (run it plain first, then under ncu).
usage: profile_kernel.py hcub <family> <nprob> [dim] | quadgk <nprob> | fixed <family> <nprob> | sampled <n> |
       hcubvec <family> <nprob> [dim] [ncomp] | bridge [dim] [log2 box budget]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import integrals_b200 as ib
import integrals_b200.workloads as WL
eng = ib.Engine(0); F = ib.FAMILY_IDS
MAXEV = int(os.environ.get("PK_MAXEVALS", "4000000"))
if os.environ.get("PK_HCUB_MODE"): eng.set_option("hcub_mode", int(os.environ["PK_HCUB_MODE"]))
kind = sys.argv[1]
reps = 2
for rep in range(reps):
    if kind == "hcub":
        fam, nprob = sys.argv[2], int(sys.argv[3]); d = int(sys.argv[4]) if len(sys.argv) > 4 else 4
        par = torch.from_numpy(np.ascontiguousarray(WL.genz_params(fam, d, nprob).T)).cuda()
        oI = torch.empty(nprob, dtype=torch.float64, device="cuda"); oN = torch.empty(nprob, dtype=torch.int64, device="cuda")
        mode = int(os.environ.get("PK_MODE", "0"))
        lbp, ubp = np.zeros(d), np.ones(d)
        if os.environ.get("PK_INF"):             # two infinite axes (the shape of BASELINE config 5)
            lbp[:2] = -np.inf; ubp[:2] = np.inf
        eng.hcubature_batch(F[fam], d, lbp, ubp, par, reltol=1e-6, abstol=1e-8, maxevals=MAXEV, out_I=oI, mode=mode,
                            out_E=torch.empty_like(oI), out_nevals=oN, out_status=torch.empty(nprob, dtype=torch.int32, device="cuda"))
        eng.synchronize()
        # the work of this launch, for scripts/executed_fp64.py (executed FP64 instructions per integrand evaluation)
        os.makedirs("gpurun_out", exist_ok=True)
        import json
        json.dump({"kind": "hcub", "family": fam, "dim": d, "nprob": nprob, "mode": mode, "maxevals": MAXEV, "evals": int(oN.sum().item()),
                   "kernel_ms": eng.last_kernel_ms}, open(f"gpurun_out/pk_hcub_{fam}_{nprob}_m{mode}.json", "w"))
    elif kind == "quadgk":
        n = int(sys.argv[2]); par = torch.from_numpy(np.ascontiguousarray(WL.c1_params(n).T)).cuda()
        oI = torch.empty(n, dtype=torch.float64, device="cuda")
        eng.quadgk_batch(F["sinxp"], np.array([-2.0]), np.array([5.0]), par, reltol=1e-10, abstol=1e-8, out_I=oI, out_E=torch.empty_like(oI),
                         out_nevals=torch.empty(n, dtype=torch.int64, device="cuda"), out_status=torch.empty(n, dtype=torch.int32, device="cuda"))
    elif kind == "fixed":
        fam, nprob = sys.argv[2], int(sys.argv[3]); x, w = np.polynomial.legendre.leggauss(64)
        par = torch.from_numpy(np.ascontiguousarray(WL.genz_params(fam, 3, nprob).T)).cuda()
        eng.fixed_rule_tensor_batch(F[fam], 3, np.zeros(3), np.ones(3), par, x, w, out=torch.empty(nprob, dtype=torch.float64, device="cuda"))
    elif kind == "sampled":
        n = int(sys.argv[2]); m = 4096
        y = torch.rand((n, m), dtype=torch.float64, device="cuda"); out = torch.empty(m, dtype=torch.float64, device="cuda")
        eng.sampled("trapezoid", y, m, n, 1, h=1.0 / (n - 1), out=out); eng.sampled("simpson", y, m, n, 1, h=1.0 / (n - 1), out=out)
    elif kind == "hcubvec":      # vector-valued HCubature: hv::hcub_vec_kernel
        fam, nprob = sys.argv[2], int(sys.argv[3]); d = int(sys.argv[4]) if len(sys.argv) > 4 else 3; nc = int(sys.argv[5]) if len(sys.argv) > 5 else 4
        par = np.asfortranarray(WL.genz_params(fam, d, nprob * nc, scale=0.25).reshape(2 * d, nc, nprob, order="F"))
        I, E, ne, st = eng.hcubature_vec_batch(F[fam], d, np.zeros(d), np.ones(d), par, reltol=1e-6, abstol=1e-8, maxevals=min(MAXEV, 400000))
        print("points", int(ne.sum()), "capped", float((st != 0).mean()))
    elif kind == "bridge":       # BatchIntegralFunction bridge (rank 3): bb::points_kernel / bb::combine_kernel, torch integrand on the device
        d = int(sys.argv[2]) if len(sys.argv) > 2 else 4; budget = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 18)
        pr = WL.genz_params("genz_gaussian", d, 1)[:, 0]
        a = torch.tensor(pr[:d], device="cuda"); u = torch.tensor(pr[d:], device="cuda")
        I, E, st = eng.integrate_batch(lambda x: torch.exp(-((a * (x - u)) ** 2).sum(dim=1)), d, np.zeros(d), np.ones(d), reltol=1e-9, abstol=0.0,
                                       max_boxes=budget, device=True)
        print("I", I, "E", E, st)
    eng.synchronize()
    print(kind, "rep", rep, "kernel ms", eng.last_kernel_ms, flush=True)
