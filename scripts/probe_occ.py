"""steady-state throughput of the pair-mode HCubature kernel vs resident warps per SM (equal-cost problems: everything capped)"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
import integrals_b200 as ib
from util_gpu import genz_params, F
eng = ib.Engine(0); eng.set_option("hcub_mode", 2)
fam = sys.argv[1] if len(sys.argv) > 1 else "genz_gaussian"
nprob = 1 << 18; MAXEV = 100000
par = torch.from_numpy(np.ascontiguousarray(genz_params(fam, 4, nprob).T)).cuda()
oI = torch.empty(nprob, dtype=torch.float64, device="cuda"); oE = torch.empty_like(oI)
oN = torch.empty(nprob, dtype=torch.int64, device="cuda"); oS = torch.empty(nprob, dtype=torch.int32, device="cuda")
for w in (4, 8, 12):
    eng.set_option("adaptive_warps_per_sm", w)
    for rep in range(2):
        eng.hcubature_batch(F[fam], 4, np.zeros(4), np.ones(4), par, reltol=1e-6, abstol=1e-8, maxevals=MAXEV, out_I=oI, out_E=oE, out_nevals=oN, out_status=oS)
        eng.synchronize(); ms = eng.last_kernel_ms
    ev = int(oN.sum().item())
    print(fam, "warps/SM limit", w, "ms", round(ms, 2), "Gevals/s", round(ev / ms / 1e6, 1), flush=True)
