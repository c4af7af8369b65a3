"""One short invocation of a hot-path kernel for ncu (run it plain first, then under ncu).
usage: profile_kernel.py hcub <family> <nprob> [dim] | quadgk <nprob> | fixed <family> <nprob> | sampled <n>"""
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
        oI = torch.empty(nprob, dtype=torch.float64, device="cuda")
        eng.hcubature_batch(F[fam], d, np.zeros(d), np.ones(d), par, reltol=1e-6, abstol=1e-8, maxevals=MAXEV, out_I=oI,
                            out_E=torch.empty_like(oI), out_nevals=torch.empty(nprob, dtype=torch.int64, device="cuda"),
                            out_status=torch.empty(nprob, dtype=torch.int32, device="cuda"))
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
    eng.synchronize()
    print(kind, "rep", rep, "kernel ms", eng.last_kernel_ms, flush=True)
