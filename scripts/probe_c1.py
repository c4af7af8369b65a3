"""C1 (QuadGK sweep) kernel time under both kernels: python scripts/probe_c1.py [nprob]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import integrals_b200 as ib
import integrals_b200.workloads as WL
eng = ib.Engine(0); F = ib.FAMILY_IDS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
par = torch.from_numpy(np.ascontiguousarray(WL.c1_params(n).T)).cuda()
oI = torch.empty(n, dtype=torch.float64, device="cuda"); oE = torch.empty_like(oI)
oN = torch.empty(n, dtype=torch.int64, device="cuda"); oS = torch.empty(n, dtype=torch.int32, device="cuda")
for mode in (1, 2, 3):
    eng.set_option("quadgk_mode", mode)
    for rep in range(4):
        eng.quadgk_batch(F["sinxp"], np.array([-2.0]), np.array([5.0]), par, reltol=1e-10, abstol=1e-8, out_I=oI, out_E=oE, out_nevals=oN, out_status=oS)
        eng.synchronize()
    ev = int(oN.sum().item())
    print("mode", mode, "kernel ms", round(eng.last_kernel_ms, 4), "Mintegrals/s", round(n / eng.last_kernel_ms / 1e3, 2), "Gevals/s", round(ev / eng.last_kernel_ms / 1e6, 2),
          "max nevals", int(oN.max().item()), flush=True)
