"""HCubature throughput probe at scale (4-D Genz families, C4 settings)."""
import json, sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
import integrals_b200 as ib
from util_gpu import genz_params, F
eng = ib.Engine(0)
if len(sys.argv) > 3: eng.set_option("hcub_mode", int(sys.argv[3]))
nprob = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 15
fams = sys.argv[2].split(",") if len(sys.argv) > 2 else ["genz_corner_peak", "genz_discontinuous", "genz_c0", "genz_gaussian", "genz_oscillatory", "genz_product_peak"]
MAXEV = int(sys.argv[4]) if len(sys.argv) > 4 else 4_000_000
res = {}
for fam in fams:
    par = torch.from_numpy(np.ascontiguousarray(genz_params(fam, 4, nprob).T)).cuda()
    oI = torch.empty(nprob, dtype=torch.float64, device="cuda"); oE = torch.empty_like(oI)
    oN = torch.empty(nprob, dtype=torch.int64, device="cuda"); oS = torch.empty(nprob, dtype=torch.int32, device="cuda")
    for rep in range(2):
        eng.hcubature_batch(F[fam], 4, np.zeros(4), np.ones(4), par, reltol=1e-6, abstol=1e-8, maxevals=MAXEV,
                            out_I=oI, out_E=oE, out_nevals=oN, out_status=oS)
        eng.synchronize()
        ms = eng.last_kernel_ms
    ev = int(oN.sum().item())
    res[fam] = dict(ms=ms, integrals_per_s=nprob / (ms * 1e-3), evals=ev, gevals_per_s=ev / (ms * 1e-3) / 1e9,
                    frac_capped=float((oS == 1).double().mean().item()), mean_evals=ev / nprob)
    print(fam, res[fam], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe_hcub.json", "w"), indent=1)
