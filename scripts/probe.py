"""First-contact GPU probe: FP64 peak, and rough timings of the four batched configs."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import integrals_b200 as ib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util_gpu import genz_params, F

eng = ib.Engine(0)
res = {}
t, ms = eng.fp64_peak(); res["fp64_peak_tflops"] = t; print("fp64 peak", t, "TFLOP/s", ms, "ms", flush=True)

def timeit(fn, reps=3):
    fn(); eng.synchronize(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        fn(); eng.synchronize()
        best = min(best, eng.last_kernel_ms)
    return best

# C2 sampled: m=4096, n = 2^18 (8 GiB) and full 2^20 (32 GiB)
for n in (1 << 16, 1 << 18, 1 << 20):
    m = 4096
    y = torch.empty((n, m), dtype=torch.float64, device="cuda")
    y.uniform_(1.0, 2.0)
    out = torch.empty(m, dtype=torch.float64, device="cuda")
    for rule in ("trapezoid", "simpson"):
        ms = timeit(lambda: eng.sampled(rule, y, m, n, 1, h=1.0 / (n - 1), out=out))
        gbs = (8.0 * m * n + 8 * m) / (ms * 1e-3) / 1e9
        res[f"sampled_{rule}_n{n}"] = dict(ms=ms, gbs=gbs); print("sampled", rule, n, ms, "ms", gbs, "GB/s", flush=True)
    del y
    torch.cuda.empty_cache()

# C3 fixed tensor rule: 2^14 problems of 64^3 nodes
x, w = np.polynomial.legendre.leggauss(64)
for fam in ("genz_oscillatory", "genz_gaussian", "genz_product_peak"):
    nprob = 1 << 14
    par = torch.from_numpy(np.ascontiguousarray(genz_params(fam, 3, nprob).T)).cuda()
    out = torch.empty(nprob, dtype=torch.float64, device="cuda")
    ms = timeit(lambda: eng.fixed_rule_tensor_batch(F[fam], 3, np.zeros(3), np.ones(3), par, x, w, out=out))
    res[f"fixed_{fam}"] = dict(ms=ms, integrals_per_s=nprob / (ms * 1e-3), evals_per_s=nprob * 64 ** 3 / (ms * 1e-3))
    print("fixed", fam, ms, "ms", nprob / (ms * 1e-3), "integrals/s", nprob * 64 ** 3 / (ms * 1e-3) / 1e12, "Tevals/s", flush=True)

# C1 quadgk: 65536 p values
n = 65536
p = torch.from_numpy((0.1 + 63.9 * np.arange(n) / (n - 1)).reshape(n, 1)).cuda()
oI = torch.empty(n, dtype=torch.float64, device="cuda"); oE = torch.empty_like(oI)
oN = torch.empty(n, dtype=torch.int64, device="cuda"); oS = torch.empty(n, dtype=torch.int32, device="cuda")
for abstol in (1e-8, 1e-12):
    ms = timeit(lambda: eng.quadgk_batch(F["sinxp"], np.array([-2.0]), np.array([5.0]), p, reltol=1e-10, abstol=abstol,
                                         out_I=oI, out_E=oE, out_nevals=oN, out_status=oS))
    ev = int(oN.sum().item())
    res[f"quadgk_abstol{abstol}"] = dict(ms=ms, integrals_per_s=n / (ms * 1e-3), evals=ev, evals_per_s=ev / (ms * 1e-3), status_max=int(oS.max().item()))
    print("quadgk", abstol, ms, "ms", n / (ms * 1e-3), "integrals/s", ev / (ms * 1e-3) / 1e9, "Gevals/s", flush=True)

# C4 hcubature: 4-D, small batches per family
for fam in ("genz_corner_peak", "genz_discontinuous", "genz_c0", "genz_gaussian", "genz_oscillatory", "genz_product_peak"):
    nprob = 4096
    par = torch.from_numpy(np.ascontiguousarray(genz_params(fam, 4, nprob).T)).cuda()
    oI = torch.empty(nprob, dtype=torch.float64, device="cuda"); oE = torch.empty_like(oI)
    oN = torch.empty(nprob, dtype=torch.int64, device="cuda"); oS = torch.empty(nprob, dtype=torch.int32, device="cuda")
    t0 = time.time()
    ms = timeit(lambda: eng.hcubature_batch(F[fam], 4, np.zeros(4), np.ones(4), par, reltol=1e-6, abstol=1e-8, maxevals=4_000_000,
                                            out_I=oI, out_E=oE, out_nevals=oN, out_status=oS), reps=1)
    ev = int(oN.sum().item())
    res[f"hcub_{fam}"] = dict(ms=ms, integrals_per_s=nprob / (ms * 1e-3), evals=ev, evals_per_s=ev / (ms * 1e-3),
                              frac_capped=float((oS == 1).double().mean().item()))
    print("hcub", fam, ms, "ms", nprob / (ms * 1e-3), "integrals/s", ev / (ms * 1e-3) / 1e9, "Gevals/s capped", res[f"hcub_{fam}"]["frac_capped"], time.time() - t0, flush=True)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe.json", "w"), indent=1)
