"""C4 (4-D Genz suite) throughput of the two refinement drivers of ib200_hcubature_batch, per family.
usage: probe_rb.py [nprob per family] [modes, e.g. 1 or 0,1] [families] [warps_per_sm]"""
import json, sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import integrals_b200 as ib
import integrals_b200.workloads as WL
eng = ib.Engine(0); F = ib.FAMILY_IDS
nprob = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
modes = [int(m) for m in (sys.argv[2] if len(sys.argv) > 2 else "0,1").split(",")]
fams = sys.argv[3].split(",") if len(sys.argv) > 3 and sys.argv[3] != "all" else WL.GENZ
if len(sys.argv) > 4: eng.set_option("adaptive_warps_per_sm", int(sys.argv[4]))
c = WL.C4; d = c["dim"]
res = {}
for fam in fams:
    par = torch.from_numpy(np.ascontiguousarray(WL.genz_params(fam, d, nprob).T)).cuda()
    oI = torch.empty(nprob, dtype=torch.float64, device="cuda"); oE = torch.empty_like(oI)
    oN = torch.empty(nprob, dtype=torch.int64, device="cuda"); oS = torch.empty(nprob, dtype=torch.int32, device="cuda")
    keep = {}
    for mode in modes:
        for rep in range(2):
            eng.hcubature_batch(F[fam], d, np.zeros(d), np.ones(d), par, reltol=c["reltol"], abstol=c["abstol"], maxevals=c["maxiters"],
                                out_I=oI, out_E=oE, out_nevals=oN, out_status=oS, mode=mode)
            eng.synchronize()
            ms = eng.last_kernel_ms
        ev = int(oN.sum().item())
        keep[mode] = (oI.clone(), oE.clone(), oS.clone())
        res[f"{fam}/mode{mode}"] = dict(ms=ms, integrals_per_s=nprob / (ms * 1e-3), evals=ev, gevals_per_s=ev / (ms * 1e-3) / 1e9,
                                        frac_capped=float((oS == 1).double().mean().item()), mean_evals=ev / nprob)
        print(fam, "mode", mode, res[f"{fam}/mode{mode}"], flush=True)
    if len(keep) == 2:
        (I0, E0, S0), (I1, E1, S1) = keep[0], keep[1]
        tol = torch.clamp(c["reltol"] * I0.abs(), min=c["abstol"])
        conv = (S0 == 0) & (S1 == 0)
        r = ((I0 - I1).abs() / tol)[conv]
        res[f"{fam}/agreement"] = dict(status_equal=float((S0 == S1).double().mean().item()), max_dI_over_tol=float(r.max().item()),
                                       frac_dI_gt_tol=float((r > 1).double().mean().item()))
        print(fam, "agreement", res[f"{fam}/agreement"], flush=True)
tot = {m: sum(v["ms"] for k, v in res.items() if k.endswith(f"mode{m}")) for m in modes}
for m in modes:
    print(f"mode {m}: {tot[m]:.1f} ms for {len(fams)} x {nprob} integrals = {len(fams) * nprob / (tot[m] * 1e-3):.0f} integrals/s")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/probe_rb_{nprob}.json", "w"), indent=1)
