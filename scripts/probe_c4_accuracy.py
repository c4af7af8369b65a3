import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import integrals_b200 as ib, integrals_b200.workloads as WL
from test_gpu_fullsize import _exact_vec
from util_gpu import F
eng = ib.Engine(0); c = WL.C4; d = 4; n = c["nprob_per_family"]
for fam in ("genz_oscillatory", "genz_product_peak", "genz_gaussian"):
    par = WL.genz_params(fam, d, n)
    I, E, ne, st = eng.hcubature_batch(F[fam], d, np.zeros(d), np.ones(d), par, reltol=c["reltol"], abstol=c["abstol"], maxevals=c["maxiters"])
    conv = st == 0; tol = np.maximum(c["abstol"], c["reltol"] * np.abs(I)); ex = _exact_vec(fam, par, d)
    r = (np.abs(I - ex) / tol)[conv]
    print(fam, "conv", conv.mean(), "ratio quantiles 50/99/99.9/max", np.quantile(r, [.5, .99, .999]), r.max(), "frac>1", (r > 1).mean(), "frac>10", (r > 10).mean(), "n>10", (r > 10).sum())
    k = np.argmax(np.where(conv, np.abs(I - ex) / tol, 0)); print("  worst:", k, I[k], ex[k], E[k], ne[k], par[:, k])
